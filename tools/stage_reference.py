"""Stages the reference's Python package for the CPU-timing arm: /root/reference/gym_PBN -> baseline/_ref/gym_PBN (git-ignored,
travels to the GPU box with the gpurun snapshot; nothing is installed, nothing is modified).  Run by __graft_entry__.build()
when /root/reference exists; a no-op elsewhere."""
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference/gym_PBN")
DST = ROOT / "baseline" / "_ref" / "gym_PBN"


def stage():
    if not SRC.is_dir():
        return False
    if DST.exists():
        shutil.rmtree(DST)
    DST.parent.mkdir(parents=True, exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.xls", "*.pkl", "*.csv"))
    return True


if __name__ == "__main__":
    print("staged" if stage() else "no /root/reference here", file=sys.stderr)
