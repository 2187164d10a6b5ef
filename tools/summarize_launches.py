"""Developer tool: per-kernel summary (count, mean duration, share of the total) of an `ncu --metrics gpu__time_duration.sum --csv`
launch list.  argv: launches.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    key = (r[4].split('(')[0][:72], r[8])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += float(r[14]) / 1e3
tot = sum(a[1] for a in agg.values())
print("%-74s %-18s %5s %10s %7s" % ("kernel", "grid", "n", "mean us", "share"))
for (k, g), (n, us) in agg.items():
    print("%-74s %-18s %5d %10.1f %6.1f%%" % (k, g, n, us / n, 100 * us / tot))
print("total %.1f us over %d launches" % (tot, len(rows)))
