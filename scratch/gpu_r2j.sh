#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_threads.py tests/test_easybytes.py -x -q -m gpu > gpurun_out/pytest_new.log 2>&1; tail -n 12 gpurun_out/pytest_new.log
timeout 900 python bench.py --no-others --no-cpu > gpurun_out/r2j_bench_pong.json 2> gpurun_out/r2j_bench_pong.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench_pong.json'))
print(d['value'], d['ms_per_step'], json.dumps(d['forward'], indent=1))
PY
tail -n 5 gpurun_out/r2j_bench_pong.err
