"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`).

usage: python scripts/summarize_launches.py <launches.csv> <out summary.csv>
"""
import collections
import csv
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    h = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr, body = rows[h], rows[h + 1:]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in body:
        name = r[kn].split("(")[0].strip()
        ns = float(r[mv].replace(",", ""))
        if name.startswith("ncb_seg_"):
            name = "ncb_seg_<k> [generated whole-segment sweeps]"
        elif "tile_kernel" in name:
            name += " [full sweep]" if ns > 1e6 else " [jump-event piece]"
        tot[name] += ns
        cnt[name] += 1
    total = sum(tot.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "share_pct", "avg_us"])
        for name, ns in tot.most_common():
            w.writerow([name, cnt[name], f"{ns / 1e6:.3f}", f"{100 * ns / total:.2f}", f"{ns / cnt[name] / 1e3:.1f}"])
    print(open(dst).read())


if __name__ == "__main__":
    main()
