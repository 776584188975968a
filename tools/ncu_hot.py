"""Top stall sites of a kernel from `ncu --page source --csv --print-source sass` output: python tools/ncu_hot.py FILE [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for r in rows[2:]:
    try:
        s = int(r[isamp])
    except (ValueError, IndexError):
        continue
    tot += s
    data.append((s, r))
print("total samples", tot, "instructions", len(data))
for s, r in sorted(data, key=lambda t: -t[0])[:n]:
    top = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{s:7d} {100*s/tot:5.1f}%  ex={r[iex]:>9}  {r[ia][-6:]}  {r[isrc][:70]:70s} {top}")
