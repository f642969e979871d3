import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gym-pbn-stac_b200"))
sys.path.insert(0, str(ROOT / "tools"))
import coop_latency
coop_latency.run(sys.argv[1], Bs=(int(sys.argv[2]),), cap=1500)
