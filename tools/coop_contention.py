"""Per-update latency of the group runner when every active warp of the GPU runs groups: B envs that never reach an
attractor (cap 4096), dealt out by the resume pass — 1184 envs = one per active warp (g = 32), 2368 = two (g = 16), ..."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
import coop_latency
coop_latency.run(sys.argv[1] if len(sys.argv) > 1 else "28_15_median", Bs=(1, 148, 296, 592, 1184, 2368, 4736, 9472, 18944), cap=4096,
                 n_cubes=int(sys.argv[2]) if len(sys.argv) > 2 else 6)
