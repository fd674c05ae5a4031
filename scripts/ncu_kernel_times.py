"""Per-kernel totals of the LAST pass in an ncu launch list (csv from --metrics gpu__time_duration.sum,...)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
kn, mn, mv, idc = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
per = {}
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    per.setdefault((int(r[idc]), r[kn].split('(')[0][:60]), {})[r[mn]] = float(r[mv].replace(',', ''))
items = sorted(per.items()); n = len(items)
tot, cnt, dr = collections.Counter(), collections.Counter(), collections.Counter()
for (i, name), m in items[n // 2:]:
    tot[name] += m.get('gpu__time_duration.sum', 0); cnt[name] += 1
    dr[name] += m.get('dram__bytes_read.sum', 0) + m.get('dram__bytes_write.sum', 0)
for name, t in tot.most_common(): print(f"{t/1e3:9.1f} us x{cnt[name]:3d}  {dr[name]/1e9:7.2f} GB  {name}")
print(f"sum {sum(tot.values())/1e3:.1f} us, dram {sum(dr.values())/1e9:.2f} GB")
if len(sys.argv) > 2:
    for (i, name), m in items[n // 2:]:
        if sys.argv[2] in name: print(i, name, m.get('gpu__time_duration.sum') / 1e3)
