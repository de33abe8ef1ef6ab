"""Aggregate an ncu gpu__time_duration launch list by kernel for ONE sampling step:
python tools/launch_agg.py gpurun_out/launches_sampler.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value')
recs = [(r[kn], float(r[mv].replace(',', ''))) for r in rows[hi + 1:] if len(r) > mv]
names = [r[0] for r in recs]
upd = [i for i, n in enumerate(names) if 'update_kernel' in n]
a, b = upd[-2] + 1, upd[-1] + 1        # a step ends with its update kernel (which also advances the step index)
agg = collections.OrderedDict()
for n, v in recs[a:b]:
    key = n.split('(')[0][:100]
    agg.setdefault(key, [0, 0.0]); agg[key][0] += 1; agg[key][1] += v
tot = sum(v for _, v in agg.values())
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f'{v/1000:9.1f} us {100*v/tot:5.1f}%  x{c:3d}  {k}')
print(f'one sampling step: {b-a} launches, sum of kernel durations {tot/1000:.1f} us')
