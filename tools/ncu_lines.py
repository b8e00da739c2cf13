"""Summarise an ncu report per CUDA source line: stall samples and executed instructions.

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N]
(reads the report here on the CPU box: ncu -i ... --page source --print-source cuda,sass --csv)
"""
import csv
import io
import os
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    path, col, agg, seen_kernel = "?", None, {}, 0
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            path = os.path.basename(r[1])
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            col = {}
            for i, n in enumerate(r):
                col.setdefault(n, i)
            continue
        if col is None or not r[0].isdigit():
            continue
        try:
            s = float(r[col["# Samples"]] or 0)
            n = float(r[col["Instructions Executed"]] or 0)
        except (ValueError, IndexError):
            continue
        key = f"{path}:{r[0]}: {r[1].strip()[:110]}"
        e = agg.setdefault(key, [0.0, 0.0])
        e[0] += s
        e[1] += n
    tot = sum(v[0] for v in agg.values()) or 1.0
    toti = sum(v[1] for v in agg.values()) or 1.0
    print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * v[0] / tot:5.1f}% samples {100 * v[1] / toti:5.1f}% inst  {k}")


if __name__ == "__main__":
    main()
