"""Issue rate of single opcode classes on this GPU (pbn_issue_peak kinds 2..5) next to the LOP3+IADD3 mix the roofline
denominator is measured with: which pipe bounds a given instruction mix.  python tools/op_peaks.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "gym-pbn-stac_b200"))
import torch  # noqa: E402

from gym_PBN.b200 import engine  # noqa: E402

sms = torch.cuda.get_device_properties(0).multi_processor_count
mhz = 1965  # the SM clock under load on this pool (bench.py samples it during its timed region)
for kind, name in ((0, "LOP3+IADD3 mix"), (2, "LOP3"), (3, "SHF"), (4, "IMAD"), (5, "IADD")):
    rate, ms = engine.issue_peak(kind, 2000)
    print(f"{name:16s} {rate:.4g} thread-ops/s = {rate / (sms * 4 * 32 * mhz * 1e6):.3f} of 1 instr/cycle/scheduler at {mhz} MHz ({ms:.2f} ms)")
