"""B200-native engine under the gym_PBN drop-in classes: ctypes binding of libpbn_b200.so (abi),
host-side network compiler front end (compiler), device-state simulator (engine), batched vector
env (vector_env) and multi-GPU helpers (dist)."""
