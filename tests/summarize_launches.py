"""Experiment helper (not a test): turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel table kept
under profiles/.   python tests/summarize_launches.py launches.csv [header line ...] > profiles/rN_launches_bench.txt"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    m = re.search(r"(k_[a-z0-9_]+)(<[^>]*?>)?", name)
    base = m.group(1) if m else name[:40]
    targs = re.match(r"k_[a-z0-9_]+<([0-9, a-z]+?)[,>]\s*(?:double|float|int|prfdd|\[|$)", name[name.find(base):]) if m else None
    lead = re.findall(r"k_[a-z0-9_]+<((?:\(?[a-z]*\)?-?[0-9]+(?:, )?)+)", name)
    tag = "<%s>" % lead[0].rstrip(", ") if lead else ""
    fn = re.search(r"(?:int )?(?:prfdd_|t_)([a-z0-9_]+)(?:<[a-z]+>)?\(", name) or re.search(r"prfdd_([a-z0-9_]+)::", name)
    vt = " f32" if re.search(r"\bfloat\b", name) else ""
    return "%s%s%s%s" % (base, tag, vt, " [%s]" % fn.group(1) if fn else "")


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    n = 0
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        t = float(r[vi].replace(",", ""))
        t = t / 1000.0 if r[ui] in ("ns", "nsecond") else t * (1000.0 if r[ui] in ("ms", "msecond") else 1.0)
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += t
        total += t; n += 1
    for h in sys.argv[2:]:
        print("# " + h)
    print("# %d launches, %.2f ms of kernel time" % (n, total / 1000.0))
    fam = OrderedDict([("SpMV family (k_spmv*, k_spmv_sell*)", 0.0), ("SEM operator (k_ax*)", 0.0), ("dense coarse product (k_dense*)", 0.0), ("vector / reduction / Krylov scalars / gather-scatter / casts", 0.0)])
    for k, (c, t) in agg.items():
        key = list(fam)[0] if k.startswith("k_spmv") else list(fam)[1] if k.startswith("k_ax") else list(fam)[2] if k.startswith("k_dense") else list(fam)[3]
        fam[key] += t
    for k, t in fam.items():
        print("# share %-62s %5.1f %%" % (k, 100.0 * t / total))
    print("# launches | total_us | share | avg_us | kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%7d | %10.1f | %5.1f%% | %8.2f | %s" % (c, t, 100.0 * t / total, t / c, k))


if __name__ == "__main__":
    main()
