"""Per CUDA source line stall samples from `ncu --page source --csv --print-source cuda,sass`: python tools/ncu_lines.py FILE [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
out = []
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
        continue
    if hdr and len(r) > 6 and r[0] not in ("", "Line No"):
        try:
            out.append((int(r[isamp]), int(r[iex]), cur_file, r[0], r[1].strip()[:110]))
        except ValueError:
            pass
tot = sum(o[0] for o in out)
print("total samples", tot)
for s, ex, f, ln, src in sorted(out, reverse=True)[:n]:
    print(f"{s:7d} {100*s/tot:5.1f}% ex={ex:>10} {f}:{ln}  {src}")
