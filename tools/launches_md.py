#!/usr/bin/env python
"""ncu launch list (csv of `ncu --metrics gpu__time_duration.sum --clock-control none --csv`) -> the per-step summary kept under
profiles/.  A step = the launches from one sample_triples_kernel (first kernel of the graph-replayed step) to the next.
    python tools/launches_md.py gpurun_out/launches.csv "header line" > profiles/<name>.md"""
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
    name = re.sub(r"\((.*)\)$", "", name)
    return name.replace("at::native::", "at::")[:84]


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if len(r) > 5]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    ls = []
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)
        ls.append((short(r[ik]), v))
    starts = [i for i, (n, _) in enumerate(ls) if n.startswith("sample_triples_kernel")]
    if len(starts) >= 2:
        step = ls[starts[-2]:starts[-1]]
    else:
        step = ls
    tot = sum(v for _, v in step)
    print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''}")
    print("# ONE graph-replayed training step (device sampler inside); per-launch times are cold-cache and serialised: compare SHARES")
    print(f"# {len(step)} launches, sum {tot:.1f} us\n#\n# in launch order: us  kernel")
    for n, v in step:
        print(f"{v:8.1f}  {n}")
    agg = {}
    for n, v in step:
        k = re.sub(r"<.*", "", n)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    print("#\n# by kernel: launches  us  share")
    for k, (c, v) in sorted(agg.items(), key=lambda t: -t[1][1]):
        print(f"{c:4d} {v:8.1f} {100 * v / tot:5.1f} %  {k}")


if __name__ == "__main__":
    main()
