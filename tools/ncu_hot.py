"""Top stall sites from an `ncu --page source --csv` dump: python tools/ncu_hot.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ci, cs, ce = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for k, r in enumerate(rows[2:]):
    try:
        data.append((float(r[cs]), k, r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
for s, k, r in sorted(data, key=lambda x: -x[0])[:N]:
    top = sorted(((float(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f'{k:5d} {100*s/tot:5.1f}% exec={r[ce]:>8} {r[ci].strip()[:70]:70s} {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f}')
