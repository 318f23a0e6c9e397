#!/bin/bash
# Forward (inference-only net): pre-split conv1 on / off
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu --no-others --steps 3 --warmup 3 > gpurun_out/r4k_bench_ps.json 2> gpurun_out/r4k_bench_ps.err
DDRL_NO_PRESPLIT_INFER=1 timeout 600 python bench.py --no-cpu --no-others --steps 3 --warmup 3 > gpurun_out/r4k_bench_nops.json 2> gpurun_out/r4k_bench_nops.err
python - <<'PY'
import json
for f in ("ps", "nops"):
    d = json.loads([l for l in open("gpurun_out/r4k_bench_%s.json" % f) if l.startswith("{")][0])
    fw = d["forward"]
    print(f, d["value"], d["ms_per_step"], "fwd", fw["value"], "u8", fw["e2e_wire"]["u8"]["value"], {k: (v["actions_per_s"], v["latency_ms"]) for k, v in fw["batch_grid"].items()})
PY
