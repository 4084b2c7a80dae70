"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r: hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    v = float(d['Metric Value'].replace(',', '')); unit = d['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    agg[re.sub(r'\(.*', '', d['Kernel Name'])][0] += 1
    agg[re.sub(r'\(.*', '', d['Kernel Name'])][1] += v
tot = sum(v[1] for v in agg.values())
print("# %s: %d launches, %.1f us total (cold-cache, serialised: compare shares)" % (sys.argv[1], sum(v[0] for v in agg.values()), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %6d launches %12.1f us %5.1f%% avg %9.1f us" % (k[:70], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
