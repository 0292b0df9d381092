"""Per-opcode and per-instruction stall summary of the LARGEST kernel section of an ncu report (source page).
    python tools/ncu_agg.py report.ncu-rep [top_n_instructions]"""
import collections
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
best = None
for a, b in zip(starts[:-1], starts[1:]):
    sect = rows[a:b]
    hi = next(i for i, r in enumerate(sect) if "Source" in r and "# Samples" in r)
    hdr = sect[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in sect[hi + 1:] if len(r) >= len(hdr) - 1]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    if best is None or tot > best[0]:
        best = (tot, hdr, ix, data, sect[0][1])
tot, hdr, ix, data, name = best
print("kernel:", name[:100], "| samples", tot)
stk = [k for k in hdr if k.startswith("stall_") and "Not" not in k]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
allst = collections.Counter()
for r in data:
    toks = r[ix["Source"]].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "")
    parts = op.split(".")
    op = parts[0] + ("." + parts[1] if len(parts) > 1 and parts[0] in ("LD", "ST", "LDS", "ATOM", "ATOMS", "MUFU", "BAR", "SYNCS", "LDG", "RED", "VOTE") else "")
    a = agg[op]
    a[0] += int(r[ix["# Samples"]]); a[1] += int(r[ix["Instructions Executed"]])
    for k in stk:
        v = r[ix[k]]
        if v not in ("", "0"):
            a[2][k[6:]] += int(v); allst[k[6:]] += int(v)
texec = sum(a[1] for a in agg.values())
print("stall reasons overall:", {k: f"{100 * v / tot:.1f}%" for k, v in allst.most_common(12)})
print("total warp instructions executed", texec)
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{op:18s} samples {a[0]:7d} {100 * a[0] / tot:5.1f}%  exec {a[1]:11d} {100 * a[1] / texec:5.1f}%  {dict(a[2].most_common(4))}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("--- top instructions")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n]:
    st = {k[6:]: int(r[ix[k]]) for k in stk if r[ix[k]] not in ("", "0")}
    print(f"{int(r[ix['# Samples']]):6d} {100 * int(r[ix['# Samples']]) / tot:5.1f}%  exec={r[ix['Instructions Executed']]:>9s}  "
          f"{r[ix['Source']].strip()[:70]:70s} {dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])}")
