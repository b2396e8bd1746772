"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # launches to skip (warm-up)
hdr, agg, n = None, collections.defaultdict(lambda: [0, 0.0]), 0
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r; ki = r.index("Kernel Name"); vi = r.index("Metric Value")
        continue
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    n += 1
    if n <= skip:
        continue
    name = re.sub(r"\(.*", "", r[ki])
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("%d launches, %.1f us total (launch list durations are cold-cache / serialised)" % (sum(v[0] for v in agg.values()), tot / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print("%-62s n=%5d %10.1f us %5.1f%%  avg %7.1f us" % (k[:62], v[0], v[1] / 1e3, 100 * v[1] / tot, v[1] / 1e3 / v[0]))
