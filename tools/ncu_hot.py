"""Top instructions of one kernel of an ncu report by stall samples (SASS view of `ncu --page source --csv`).
   python tools/ncu_hot.py report.ncu-rep <kernel index> [top]"""
import csv, io, subprocess, sys

rep, kidx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
secs, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = []
        secs.append(cur)
    elif cur is not None:
        cur.append(line)
rows = list(csv.reader(io.StringIO("\n".join(secs[kidx]))))
hdr, rows = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
S, I = ix["# Samples"], ix["Instructions Executed"]
tot = sum(int(r[S]) for r in rows)
toti = sum(int(r[I]) for r in rows)
print(f"kernel section {kidx}: {len(rows)} instructions, {tot} samples, {toti} warp-instructions executed")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in rows) for h in stalls}
print("stall reasons:", ", ".join(f"{k[6:]} {v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(rows)), key=lambda k: -int(rows[k][S]))[:top]
for k in sorted(order):
    r = rows[k]
    why = max(stalls, key=lambda h: int(r[ix[h]]))
    print(f"{k:5d} {int(r[S]) * 100.0 / tot:5.1f}% exec {int(r[I]):9d} thr {r[ix['Avg. Threads Executed']]:>4s} {why[6:]:14s} {r[ix['Source']].strip()[:90]}")
