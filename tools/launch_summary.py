"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for one training
step (the launches between `patchify` and `adamw`), plus per-launch GEMM durations with their template args."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = []
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        if unit == "ns":
            t /= 1e3
        elif unit == "ms":
            t *= 1e3
        per.append((row["Kernel Name"], t))
    return per


def short(name):
    n = re.sub(r"\(.*", "", name)
    return n.replace("micgemm::", "").replace("<unnamed>::", "").replace("void ", "")


def main(path, which=0, out=None):
    per = load(path)
    starts = [i for i, (n, _) in enumerate(per) if "patchify" in n]
    ends = [i for i, (n, _) in enumerate(per) if "adamw" in n]
    # the capture window (-s / -c) starts anywhere: take the which-th patchify that has an adamw after it
    pairs = [(i, min(e for e in ends if e > i)) for i in starts if any(e > i for e in ends)]
    a, b = pairs[which][0], pairs[which][1] + 1
    step = per[a:b]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for n, t in step:
        k = short(n)
        agg[k][0] += 1
        agg[k][1] += t
        tot += t
    lines = [f"{path}: step #{which}: {len(step)} launches, {tot:.0f} us serialised"]
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"{t:10.0f} us {100 * t / tot:5.1f}%  n={c:4d}  {k[:110]}")
    txt = "\n".join(lines)
    print(txt)
    if out:
        open(out, "w").write(txt + "\n")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, sys.argv[3] if len(sys.argv) > 3 else None)
