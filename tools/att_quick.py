"""Prints the step-until-attractor lines of tools/bench_configs.py compactly (kernel experiments)."""
import json
import subprocess
import sys

out = subprocess.run([sys.executable, "tools/bench_configs.py"], capture_output=True, text=True).stdout
for line in out.splitlines():
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    c = d["config"]
    if "wall clock" in c or "rollout" in c or "synthetic" in c:
        continue
    print(f"{c[:34]:34s} ... {c[-38:]:38s} {d['env_steps_per_s']:.4g} env-steps/s  {d.get('ms_per_step', 0):.3f} ms")
