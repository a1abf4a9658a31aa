"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, re, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]); name = re.sub(r"void |<unnamed>::|at::native::|\(anonymous namespace\)::", "", name)[:84]
    agg[name][0] += 1; agg[name][1] += float(r[14].replace(",", "")) / 1e6
tot = sum(v[1] for v in agg.values())
print("total kernel time %.2f ms over %d launches" % (tot, len(rows)))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%7.3f ms %5.1f%% %5d  %s" % (v[1], 100 * v[1] / tot, v[0], k))
