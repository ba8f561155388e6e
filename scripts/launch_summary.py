"""Summarise an ncu launch-list csv (gpu__time_duration per kernel)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict(); seq = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0].startswith('#'): continue
    name = r[ix['Kernel Name']].split('(')[0][-44:]
    # raw ncu csv ('Metric Value' / 'Metric Unit') or the trimmed copy under profiles/ (last two columns)
    v = float(r[ix['Metric Value']] if 'Metric Value' in ix else r[-2]); u = r[ix['Metric Unit']] if 'Metric Unit' in ix else r[-1]
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v * 1e6 if u == 's' else v
    seq.append((name, v)); a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k, (c, t) in agg.items(): print(f"{k:46s} n={c:4d} total={t:10.1f}us avg={t/c:9.1f}us")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
print("last launches:", [(a[-26:], round(b, 1)) for a, b in seq[-n:]])
