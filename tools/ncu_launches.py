"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

usage: python tools/ncu_launches.py <launches.csv> [--top-regex N]
Prints kernel, launches, total us, share of the listed time; with --top-regex prints a regex that
matches the N kernels with the largest total (for a follow-up `ncu --set full -k regex:...`).
"""
import csv
import re
import sys


def load(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = {n: i for i, n in enumerate(r)}
            continue
        if len(r) < len(hdr):
            continue
        if r[hdr["Metric Name"]] != "gpu__time_duration.sum":
            continue
        unit = r[hdr["Metric Unit"]]
        v = float(r[hdr["Metric Value"]].replace(",", ""))
        us = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)
        rows.append((r[hdr["Kernel Name"]], us))
    return rows


def short(name):
    m = re.match(r"(?:void\s+)?(?:[\w:]+::)?(\w+)", name)
    return m.group(1) if m else name


def main():
    rows = load(sys.argv[1])
    agg = {}
    for k, us in rows:
        e = agg.setdefault(short(k), [0, 0.0])
        e[0] += 1
        e[1] += us
    tot = sum(v[1] for v in agg.values()) or 1.0
    ranked = sorted(agg.items(), key=lambda kv: -kv[1][1])
    if len(sys.argv) > 2 and sys.argv[2] == "--top-regex":
        n = int(sys.argv[3]) if len(sys.argv) > 3 else 5
        print("(" + "|".join(k for k, _ in ranked[:n]) + ")")
        return
    print(f"{len(rows)} launches, {tot / 1e3:.3f} ms listed (cold-cache, serialised: compare shares)")
    print(f"{'kernel':32s} {'launches':>8s} {'total_us':>12s} {'us/launch':>10s} {'share':>7s}")
    for k, (n, us) in ranked:
        print(f"{k:32s} {n:8d} {us:12.1f} {us / n:10.1f} {100 * us / tot:6.2f}%")


if __name__ == "__main__":
    main()
