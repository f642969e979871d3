"""Markdown summary of an ncu report of the step-until-attractor kernels (run here, no GPU needed).
    python tools/ncu_att_summary.py gpurun_out/r02c_att.ncu-rep profiles/r02_att_ncu_summary.md "<command that was profiled>"
Per captured launch: key metrics of the raw page, the stall-reason mix and the hottest instructions of the source page
(SASS view; the kernels are compiled with -lineinfo, `ncu -i <report> --page source` shows the CUDA lines)."""
import csv
import io
import subprocess
import sys

rep, out_md, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "registers/thread"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / instruction"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "of which bank-conflict replays"),
        ("sm__cycles_elapsed.max", "cycles elapsed (max)"), ("sm__cycles_active.avg", "cycles active (avg over SMs)"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written")]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
secs, cur = [], None
for line in src.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = []
        secs.append(cur)
    elif cur is not None:
        cur.append(line)
secs = secs[::2] if len(secs) == 2 * len(data) else secs  # the page lists every kernel twice (SASS, source)
md = [f"# ncu summary — step-until-attractor kernels (round 2)\n", f"Command: `{cmd}`  \nReport: `{rep}` (`ncu --set full --clock-control none --import-source on`)\n"]
for k, r in enumerate(data):
    name = r[ix["Kernel Name"]].split("(")[0]
    md.append(f"\n## launch {k}: `{name}`\n\n| metric | value |\n|---|---|")
    for key, label in KEYS:
        if key in ix:
            md.append(f"| {label} | {r[ix[key]]} {units[ix[key]]} |")
    if k < len(secs):
        srows = list(csv.reader(io.StringIO("\n".join(secs[k]))))
        sh, srows = srows[0], srows[1:]
        six = {h: i for i, h in enumerate(sh)}
        S = six["# Samples"]
        tot = sum(int(x[S]) for x in srows) or 1
        stalls = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
        agg = sorted(((sum(int(x[six[h]]) for x in srows), h[6:]) for h in stalls), reverse=True)[:7]
        md.append("\nStall samples: " + ", ".join(f"{n} {100 * v // tot}%" for v, n in agg))
        md.append("\nHottest instructions (share of stall samples, executions, avg active threads, top stall, SASS):\n\n```")
        order = sorted(range(len(srows)), key=lambda q: -int(srows[q][S]))[:14]
        for q in sorted(order):
            x = srows[q]
            why = max(stalls, key=lambda h: int(x[six[h]]))
            md.append(f"{100.0 * int(x[S]) / tot:5.1f}%  x{int(x[six['Instructions Executed']]):>9d}  thr {x[six['Avg. Threads Executed']]:>3s}  {why[6:]:15s} {x[six['Source']].strip()[:80]}")
        md.append("```")
open(out_md, "w").write("\n".join(md) + "\n")
print("wrote", out_md)
