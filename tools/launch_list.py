"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` log into a per-kernel launch list (count, total, average, share)."""
import csv
import sys
from collections import OrderedDict


def main(path, header=""):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(r[iu], 1e-6)
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    if header:
        print(header)
    print(f"# {sum(a[0] for a in agg.values())} launches, {tot:.1f} ms of kernel time")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:100]:102s} n={n:4d} total_ms={ms:10.3f} avg_ms={ms / n:10.4f} share={100 * ms / tot:6.2f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
