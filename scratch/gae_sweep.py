#!/usr/bin/env python
"""BASELINE C3 sweep: GAE schedules over N in {1k,4k,16k,64k} x T in {128,512,2048}; prints JSON lines.
usage: python scratch/gae_sweep.py [out.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ddrl4nav_b200 import kernels  # noqa: E402

dev = torch.device("cuda", 0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []
for N in (1024, 4096, 16384, 65536):
    for T in (128, 512, 2048):
        gen = torch.Generator(device=dev).manual_seed(N + T)
        v = torch.randn(T + 1, 1, N, device=dev, generator=gen)
        r = torch.randn(T, 1, N, device=dev, generator=gen)
        d = (torch.rand(T, 1, N, device=dev, generator=gen) < 0.02).to(torch.uint8)
        ret1, adv1 = kernels.gae(v, r, d, [0.99], 0.95, 1)
        row = {"N": N, "T": T, "bytes": 17 * T * N}
        for algo in (0, 1, 2, 4, 5, 7):
            ret, adv = kernels.gae(v, r, d, [0.99], 0.95, algo)
            err = float((adv - adv1).abs().max() / adv1.abs().max())
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                kernels.gae(v, r, d, [0.99], 0.95, algo)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            gbs = 17 * T * N / (ms * 1e-3) / 1e9
            row["algo%d" % algo] = {"us": round(ms * 1e3, 1), "GBps": round(gbs, 1), "frac": round(gbs / peak, 3), "err": err}
        rows.append(row)
        print(json.dumps(row), flush=True)
if len(sys.argv) > 1:
    json.dump({"peak_hbm_gbs": peak, "note": "median of 7, L2 flushed (256 MB write) before every launch, CUDA events incl. launch",
               "rows": rows}, open(sys.argv[1], "w"), indent=1)
