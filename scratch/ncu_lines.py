#!/usr/bin/env python
"""Per-CUDA-line warp-stall samples from an ncu report (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; kern = None
agg = collections.defaultdict(lambda: [0, 0, ""]); stall = collections.defaultdict(lambda: collections.Counter())
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": kern = r[1][:60]; continue
    if r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed"); st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]; continue
    if hdr is None or len(r) <= si: continue
    if r[2] != "-": continue            # SASS rows repeat under their line; line rows have '-' address
    try: n = int(r[si]); ne = int(r[ie])
    except ValueError: continue
    key = (kern, cur_file, int(r[0]))
    agg[key][0] += n; agg[key][1] += ne; agg[key][2] = r[1].strip()[:110]
    for i in st:
        try: v = int(r[i])
        except ValueError: v = 0
        if v: stall[key][hdr[i][6:]] += v
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
for key, (n, ne, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% %7d smp %9d inst  %s:%d  %s   [%s]" % (100.0 * n / tot, n, ne, key[1], key[2], src, " ".join("%s=%d" % kv for kv in stall[key].most_common(3))))
# role regions of tc2.cu (line ranges of the kernel body)
if any(k[1] == "tc2.cu" for k in agg):
    regions = [("setup", 98, 155), ("producer", 156, 233), ("mma", 234, 292), ("splitter", 293, 358), ("epi-drain", 359, 407), ("epi-store", 408, 549), ("tail", 550, 560)]
    print("region        samples   share   warp-instructions")
    for nm, a, b in regions:
        sm = sum(v[0] for k, v in agg.items() if k[1] == "tc2.cu" and a <= k[2] <= b)
        ins = sum(v[1] for k, v in agg.items() if k[1] == "tc2.cu" and a <= k[2] <= b)
        print("%-12s %8d  %5.1f%%  %12d" % (nm, sm, 100.0 * sm / tot, ins))
    sm = sum(v[0] for k, v in agg.items() if k[1] != "tc2.cu"); ins = sum(v[1] for k, v in agg.items() if k[1] != "tc2.cu")
    print("%-12s %8d  %5.1f%%  %12d   (inlined helpers: tc_ptx.cuh etc., all roles)" % ("other files", sm, 100.0 * sm / tot, ins))
