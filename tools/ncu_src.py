"""Summarise an ncu --page source --csv dump: top stalled SASS instructions with stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
def f(x):
    try: return float(x)
    except Exception: return 0.0
s = ci["Warp Stall Sampling (All Samples)"]
tot = sum(f(r[s]) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
idx = {id(r): i for i, r in enumerate(data)}
for r in sorted(data, key=lambda r: -f(r[s]))[:n]:
    top = sorted(((f(r[ci[h]]), h[6:]) for h in stalls), reverse=True)[:3]
    print(f"{idx[id(r)]:5d} {f(r[s]):7.0f} {100*f(r[s])/tot:5.1f}% exec={r[ci['Instructions Executed']]:>9s} "
          f"{r[ci['Source']].strip()[:70]:70s} " + " ".join(f"{h}={v:.0f}" for v, h in top if v > 0))
