"""Condense an `ncu --set full` report into the few numbers DESIGN.md / bench.py quote.

usage: python tools/ncu_summary.py <report.ncu-rep> [--traffic profiles/traffic.json]
Prints one block per profiled launch (duration, DRAM bytes, throughput percentages, occupancy,
registers, top warp-stall reasons) and, with --traffic, merges `dram read + write bytes per launch`
into a JSON map keyed by the profile name bench.py uses (k_seg_cell -> seg_cell).
"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
    ("launch__registers_per_thread", "registers"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_static", "smem static"),
    ("launch__shared_mem_per_block_dynamic", "smem dynamic"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]

TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    m = re.match(r"(?:void\s+)?(?:[\w:]+::)?(\w+)(<[^(]*>)?", name)
    if not m:
        return name, name
    tmpl = m.group(2) or ""
    first = re.match(r"<\s*(?:[\w:]+::)?(\w+)", tmpl)
    return m.group(1), m.group(1) + ("<" + first.group(1) + ">" if first else "")


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    stall_cols = [(i, n) for i, n in enumerate(hdr)
                  if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")
                  and "not_issued" not in n]
    traffic = {}
    for r in rows[2:]:
        base, disp = short(r[col["Kernel Name"]])
        print(f"== {disp}")
        vals = {}
        for key, label in KEEP:
            if key in col:
                v, u = r[col[key]], units[col[key]]
                vals[key] = (v, u)
                print(f"   {label:26s} {v} {u}")
        st = []
        for i, n in stall_cols:
            try:
                st.append((float(r[i]), n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
        st.sort(reverse=True)
        print("   top stalls (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in st[:5]))
        try:
            rd = float(vals["dram__bytes_read.sum"][0]) * TO_BYTES.get(vals["dram__bytes_read.sum"][1], 1.0)
            wr = float(vals["dram__bytes_write.sum"][0]) * TO_BYTES.get(vals["dram__bytes_write.sum"][1], 1.0)
            name = base[2:] if base.startswith("k_") else base
            if base not in ("k_compact_scatter", "k_compact_count", "k_excl_scan"):
                traffic[name] = rd + wr
        except (KeyError, ValueError):
            pass
    if "--traffic" in sys.argv:
        path = sys.argv[sys.argv.index("--traffic") + 1]
        try:
            cur = json.load(open(path))
        except Exception:
            cur = {}
        cur.update(traffic)
        json.dump(cur, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
