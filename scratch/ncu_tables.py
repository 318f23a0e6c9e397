#!/usr/bin/env python
"""Turns the ncu CSV logs a GPU session leaves in gpurun_out/ into the markdown / csv summaries committed under profiles/.

  python scratch/ncu_tables.py hbm  OUT.md  metrics1.csv [metrics2.csv ...]   # dram bytes + duration per kernel class
  python scratch/ncu_tables.py tc   OUT.md  raw.csv                            # --set full --page raw: one row per launch
  python scratch/ncu_tables.py list OUT.csv launches.csv                       # gpu__time_duration launch list -> share per kernel
"""
import collections
import csv
import re
import sys


def rows_of(path):
    rows = list(csv.reader(open(path, newline="")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    return rows[hi], rows[hi + 1:]


def short(name):
    name = name.replace("void ", "").replace("ddrl::", "")
    m = re.match(r"([A-Za-z0-9_]+(<[^>]*>)?)", name)
    return m.group(1) if m else name


def hbm(out, files):
    per = collections.defaultdict(dict)
    for f in files:
        hdr, data = rows_of(f)
        ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
        for r in data:
            if len(r) <= iv:
                continue
            d = per[(f, r[iid])]
            d["name"] = short(r[ik])
            try:
                d[r[im]] = float(r[iv].replace(",", ""))
            except ValueError:
                pass
    by = collections.defaultdict(list)
    for d in per.values():
        if "gpu__time_duration.sum" in d:
            by[d["name"]].append(d)
    with open(out, "w") as fh:
        fh.write("| kernel | launches | largest launch: us | DRAM MB (read+write) | DRAM GB/s | dram % of peak | mean us | mean GB/s |\n")
        fh.write("|---|---|---|---|---|---|---|---|\n")
        for n in sorted(by):
            vs = by[n]
            big = max(vs, key=lambda v: v["gpu__time_duration.sum"])
            byt = lambda v: v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0)
            t, b = big["gpu__time_duration.sum"], byt(big)
            mt = sum(v["gpu__time_duration.sum"] for v in vs) / len(vs)
            mb = sum(byt(v) for v in vs) / len(vs)
            fh.write("| `%s` | %d | %.1f | %.2f | %.0f | %.1f | %.1f | %.0f |\n" % (
                n, len(vs), t / 1e3, b / 1e6, b / t if t else 0, big.get("dram__throughput.avg.pct_of_peak_sustained_elapsed", 0.0),
                mt / 1e3, mb / mt if mt else 0))


UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "Kbyte/block": 1e3, "byte/block": 1.0, "Mbyte/block": 1e6}


def tc(out, path):
    """--page raw --csv of an `ncu --set full` capture: one row per launch.  The raw page scales every column to its own
    unit (second header row): durations are normalised to us, byte counts to bytes."""
    hdr, data = rows_of(path)
    units, data = data[0], data[1:]

    def col(name):
        return [i for i, h in enumerate(hdr) if h == name or h.endswith("." + name)][0]

    def val(r, name):
        i = col(name)
        try:
            return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        except (ValueError, IndexError):
            return float("nan")
    cols = [("us", "gpu__time_duration.sum", 1), ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
            ("SM busy %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1), ("issue slots busy %", "sm__inst_issued.avg.pct_of_peak_sustained_active", 1),
            ("DRAM read MB", "dram__bytes_read.sum", 1e-6), ("DRAM write MB", "dram__bytes_write.sum", 1e-6),
            ("L2->SM MB", "l1tex__m_xbar2l1tex_read_bytes.sum", 1e-6), ("L2 hit %", "lts__t_sector_hit_rate.pct", 1),
            ("grid", "launch__grid_size", 1), ("regs", "launch__registers_per_thread", 1), ("smem KB", "launch__shared_mem_per_block_dynamic", 1e-3)]
    cols = [c for c in cols if any(h == c[1] or h.endswith("." + c[1]) for h in hdr)]
    ki = col("Kernel Name")
    tot_b = tot_us = 0.0
    with open(out, "w") as fh:
        fh.write("| # | kernel | " + " | ".join(c[0] for c in cols) + " | L2->SM TB/s | DRAM TB/s |\n")
        fh.write("|---|---|" + "---|" * (len(cols) + 2) + "\n")
        for i, r in enumerate(data):
            vals = {c[0]: val(r, c[1]) * c[2] for c in cols}
            us = vals["us"]
            dram = vals["DRAM read MB"] + vals["DRAM write MB"]
            tot_b += dram
            tot_us += us
            fh.write("| %d | `%s` | " % (i, short(r[ki])) + " | ".join("%.1f" % vals[c[0]] for c in cols) +
                     " | %.2f | %.2f |\n" % (vals["L2->SM MB"] / us, dram / us))
        fh.write("\n%d launches: %.1f us in total (serialised, cold cache), DRAM read+write %.1f MB in total = %.1f MB per launch\n" %
                 (len(data), tot_us, tot_b, tot_b / max(len(data), 1)))


def launch_list(out, path):
    hdr, data = rows_of(path)
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    im = hdr.index("Metric Name")
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) > iv and r[im] == "gpu__time_duration.sum":
            t = tot[short(r[ik])]
            t[0] += 1
            t[1] += float(r[iv].replace(",", ""))
    s = sum(v[1] for v in tot.values())
    with open(out, "w") as fh:
        fh.write("kernel,launches,total_us,share\n")
        for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            fh.write("%s,%d,%.1f,%.4f\n" % (n, c, t / 1e3, t / s))


if __name__ == "__main__":
    mode, out = sys.argv[1], sys.argv[2]
    {"hbm": lambda: hbm(out, sys.argv[3:]), "tc": lambda: tc(out, sys.argv[3]), "list": lambda: launch_list(out, sys.argv[3])}[mode]()
