"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python scripts/launch_summary.py gpurun_out/<tag>/launches.csv [steps]"""
import collections
import csv
import re
import sys


def main(path, steps=1, top=60):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        name = row["Kernel Name"]
        name = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        name = re.sub(r"^void ", "", name)
        m = re.match(r"at::native::(\w+)<(?:at::native::)?([\w:]+)", name)
        name = ("at::%s<%s>" % (m.group(1), m.group(2)) if m else re.sub(r"[(<].*", "", name))[:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    print("%d launches, %.1f us total (%d step(s): %.1f us/step, cold-cache serialised ncu times)" % (n, tot, steps, tot / steps))
    print("%12s %6s %6s  %s" % ("us/step", "share", "n/step", "kernel"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%12.1f %5.1f%% %6.1f  %s" % (v[1] / steps, 100 * v[1] / tot, v[0] / steps, k))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
