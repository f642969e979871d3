"""CPU: the C-ABI library builds, loads, and exports every symbol include/pbn_b200.h declares."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    import importlib.util

    spec = importlib.util.spec_from_file_location("pbn_build", ROOT / "gym-pbn-stac_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib_path = mod.build()
    header = (ROOT / "include" / "pbn_b200.h").read_text()
    declared = set(re.findall(r"\b(pbn_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 15
    lib = ctypes.CDLL(str(lib_path))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in pbn_b200.h but not exported"
    lib.pbn_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.pbn_version()


def test_binding_table_matches_header():
    from gym_PBN.b200 import abi

    header = (ROOT / "include" / "pbn_b200.h").read_text()
    declared = set(re.findall(r"\b(pbn_[a-z_0-9]+)\s*\(", header))
    assert declared == set(abi.EXPORTS)
    abi.lib()  # loads and binds all of them
