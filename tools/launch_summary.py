"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`).
usage: python tools/launch_summary.py X.csv [skip_first_n_launches]"""
import csv, collections, sys, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
gi, bi = hdr.index('Grid Size'), hdr.index('Block Size')
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
d = collections.OrderedDict()
for r in rows[1 + skip:]:
    name = re.sub(r'\(.*', '', re.sub(r'<unnamed>::', '', r[ki].replace('void ', '')))
    key = name + ' grid ' + r[gi]
    d.setdefault(key, []).append(float(r[vi].replace(',', '')))
tot = sum(sum(v) for v in d.values())
print("total %.3f ms over %d launches" % (tot / 1e6, sum(len(v) for v in d.values())))
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print("%-90s n=%4d avg=%9.1f us  share=%5.1f%%" % (k[:90], len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
