"""Print the SASS of a kernel from an `ncu --page source --csv` dump with executed counts (hot loop first)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iA, iS, iE, iT = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed')
iSm = hdr.index('# Samples')
tot = sum(int(r[iE]) for r in rows[2:] if len(r) > iE)
print('total warp instr', tot)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
for n, r in enumerate(rows[2:]):
    e = int(r[iE])
    if e / tot >= thr:
        print(f"{n:4d} {e:10d} {e/tot*100:5.2f}% thr={r[iT]:>5s} smp={r[iSm]:>5s}  {r[iS].strip()}")
