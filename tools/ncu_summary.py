"""Summarise an ncu report of the SSD kernel into profiles/ (run here, no GPU needed).

    python tools/ncu_summary.py gpurun_out/prof_ssd.ncu-rep --iters-per-launch 10066329600 --tag r01

Writes profiles/<tag>_ssd_ncu_summary.md (key metrics per captured launch, SASS opcode mix) and
profiles/ssd_inst_per_iter.json (thread-level instructions per SSD iteration + DRAM bytes per launch), the
figures bench.py uses for the issue-rate roofline."""
import argparse
import collections
import csv
import io
import json
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_shared_atom.sum",
        "sm__cycles_active.avg"]


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v) * scale


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--iters-per-launch", type=float, required=True)
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--command", default="python bench.py --steps 2 --warmup 3")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = [f"# ncu summary — SSD kernel ({args.tag})", "",
           f"Report: `{args.report}` captured with `ncu --set full --clock-control none --import-source on -k regex:k_ssd` "
           f"around `{args.command}`; {args.iters_per_launch:.4g} SSD iterations per launch.", ""]
    last = None
    for r in data:
        out.append(f"## {r[hdr.index('Kernel Name')][:90]}")
        out.append("")
        out.append("| metric | value | unit |")
        out.append("|---|---|---|")
        vals = {}
        for k in KEYS:
            if k in hdr:
                vals[k] = (r[hdr.index(k)], units[hdr.index(k)])
                out.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        wi = float(vals["smsp__inst_executed.sum"][0])
        ti = float(vals["smsp__thread_inst_executed.sum"][0]) if "smsp__thread_inst_executed.sum" in vals else wi * float(
            vals["smsp__thread_inst_executed_per_inst_executed.ratio"][0])
        dram = to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])
        out += ["", f"* warp instructions per warp-iteration: **{wi / (args.iters_per_launch / 32):.1f}**",
                f"* thread instructions per SSD iteration: **{ti / args.iters_per_launch:.1f}**",
                f"* DRAM bytes per launch (read+write): {dram:.4g}", ""]
        if wi != wi:  # a launch whose counters ncu could not collect (nan): keep the previous one
            continue
        last = {"thread_inst_per_iter": ti / args.iters_per_launch, "warp_inst_per_warp_iter": wi / (args.iters_per_launch / 32),
                "dram_bytes_per_launch": dram, "issue_active_pct": float(vals["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
                "source": f"profiles/{args.tag}_ssd_ncu_summary.md (ncu --set full, {args.command})"}
    # opcode mix of the first captured launch
    sass = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv", "--print-source", "sass"],
                          capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(sass)))
    ops = collections.Counter()
    for r in srows[2:]:
        if r and r[0] == "Kernel Name":
            break
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1])
        if m:
            ops[m.group(2).split(".")[0]] += float(r[5])
    per = args.iters_per_launch / 32
    out += ["## SASS opcode mix (executed warp instructions per warp-iteration, first launch)", "", "| opcode | per warp-iteration |", "|---|---|"]
    for op, c in ops.most_common(24):
        out.append(f"| {op} | {c / per:.2f} |")
    (ROOT / "profiles").mkdir(exist_ok=True)
    (ROOT / "profiles" / f"{args.tag}_ssd_ncu_summary.md").write_text("\n".join(out) + "\n")
    (ROOT / "profiles" / "ssd_inst_per_iter.json").write_text(json.dumps(last, indent=1) + "\n")
    print("\n".join(out[:40]))


if __name__ == "__main__":
    main()
