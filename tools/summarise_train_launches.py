"""Per-kernel table of ONE whole training step out of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv of
tests/train_step_timing.py with SPM_TIMING_ONLY_WHOLE=1): the launches between the last two optimiser steps.
    python tools/summarise_train_launches.py gpurun_out/launches.csv "header text" > profiles/...txt"""
import collections
import csv
import re
import sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
names = [r["Kernel Name"] for r in rows]
bwd = [i for i, n in enumerate(names) if "vit_attention_bwd" in n]
adam = [i for i, n in enumerate(names) if "adam_kernel" in n]
start, end = max(a for a in adam if a < bwd[-12]) + 1, max(adam) + 1
agg = collections.defaultdict(lambda: [0, 0.0])
for row in rows[start:end]:
    n = re.sub(r"\(.*", "", row["Kernel Name"])
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"spm::\(anonymous namespace\)::|spm::|<unnamed>::", "", n)
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print("%d launches, %.1f ms of kernel time (cold-cache, serialised by ncu)\n" % (end - start, tot / 1e3))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print("%6d %10.1f us %5.1f%%  %s" % (c, t, 100 * t / tot, n[:120]))
