"""Per-source-line totals (stall samples, warp instructions executed) from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
cur = None
out = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed")
        continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        try:
            out.append((int(r[i_s]), int(r[i_i]), cur, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
ts = sum(o[0] for o in out); ti = sum(o[1] for o in out)
print(f"total samples {ts}, warp instructions {ti}")
for s, i, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100*s/ts:5.1f}% smp {100*i/ti:5.1f}% inst  {f}:{ln}  {src}")
