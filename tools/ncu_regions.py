"""Measurement aid: executed-instruction profile of one kernel from an `ncu --page source --csv` export.
Prints runs of consecutive SASS instructions with similar execution counts (code regions) and their share of the total."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) >= len(rows[1]) - 2 and r[0].startswith("0x") or (len(r)>5 and r[0][:1].isdigit())]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
out = [(r[ia], r[isrc], float(r[iex] or 0), float(r[ismp] or 0)) for r in data]
tot = sum(o[2] for o in out); tsm = sum(o[3] for o in out)
print(f'total warp-instructions {tot/1e9:.3f} G, samples {tsm:.0f}')
prev = None; start = 0; s = 0; smp = 0
def flush(k):
    print(f"{out[start][0][-5:]}..{out[k-1][0][-5:]}  n={k-start:4d}  exec/instr={prev*1e5/1e6:8.2f}M  share={s/tot*100:5.1f}%  samples={smp/tsm*100:5.1f}%  first: {out[start][1][:60]}")
for k, (a, src, ex, sm) in enumerate(out):
    key = round(ex / 1e5)
    if prev is None:
        prev = key
    if abs(key - prev) > max(2, 0.08 * prev):
        flush(k); prev = key; start = k; s = 0; smp = 0
    s += ex; smp += sm
flush(len(out))
