"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_summary.py LIST.csv [--last-step KERNEL_SUBSTRING] [--title TEXT]

--last-step: keep only the launches after the second-to-last occurrence of a kernel whose name contains the substring, up to
and including its last occurrence (one whole training step when given the optimizer kernel, one NMS call when given the
candidates kernel ...).
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    last = sys.argv[sys.argv.index("--last-step") + 1] if "--last-step" in sys.argv else None
    first = sys.argv[sys.argv.index("--from-last") + 1] if "--from-last" in sys.argv else None
    title = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else path
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    names = [r["Kernel Name"] for r in rows]
    if last:
        idx = [i for i, n in enumerate(names) if last in n]
        rows = rows[idx[-2] + 1: idx[-1] + 1]
    if first:
        idx = [i for i, n in enumerate(names) if first in n]
        rows = rows[idx[-1]:]
    agg = collections.OrderedDict()
    for r in rows:
        v = float(r["Metric Value"])
        v = v / 1e3 if r["Metric Unit"] == "ns" else v
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("y3::<unnamed>::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {title}")
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold-cache launches: compare shares, not absolutes)")
    print(f"{'kernel':74s} launches   total_us   share")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:74]:74s} {c:8d} {t:10.1f} {t / tot:7.3f}")
    print(f"{'total':74s} {sum(v[0] for v in agg.values()):8d} {tot:10.1f}")


if __name__ == "__main__":
    main()
