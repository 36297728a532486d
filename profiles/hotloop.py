"""Hot loop of a kernel from an `ncu --page source --csv` export: instructions executed >= 50 % of the maximum, with samples and the top stall reasons.

    python profiles/hotloop.py gpurun_out/r02e_source_2.csv [kernel index]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
# the file holds one table per profiled kernel, each introduced by a "Kernel Name" row
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
a = starts[which]
b = starts[which + 1] if which + 1 < len(starts) else len(rows)
hdr = rows[a + 1]
body = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
ex = [float(r[ix["Instructions Executed"]] or 0) for r in body]
mx = max(ex)
tot_samples = sum(float(r[ix["# Samples"]] or 0) for r in body)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"{rows[a][1]}; total samples {tot_samples:.0f}; loop max executed {mx:.0f}")
n = 0
counts = {}
for r, e in zip(body, ex):
    if e < 0.5 * mx:
        continue
    n += 1
    op = r[ix["Source"]].split()[0] if not r[ix["Source"]].startswith("@") else r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    counts[op] = counts.get(op, 0) + e / mx
    st = sorted(((float(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{r[ix['Address']][-5:]} {float(r[ix['# Samples']] or 0):7.0f} {100 * float(r[ix['# Samples']] or 0) / tot_samples:4.1f}%  {r[ix['Source']][:70]:70s} " + " ".join(f"{nme}={v:.0f}" for v, nme in st if v > 0))
print(f"{n} instructions in the loop; per-iteration mix: " + ", ".join(f"{k}={v:.1f}" for k, v in sorted(counts.items(), key=lambda kv: -kv[1])))
