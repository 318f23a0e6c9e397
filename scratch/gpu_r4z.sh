#!/bin/bash
# evidence pass of the round: full GPU test suite, default bench line, reference arm, ncu launch list, ncu --set full of the
# tc3 kernels, ncu DRAM metrics of the HBM-bound kernels
set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r4z_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4z_pytest_gpu.log; tail -n 6 gpurun_out/r4z_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r4z_bench.json 2> gpurun_out/r4z_bench.err; head -c 400 gpurun_out/r4z_bench.json; echo; tail -n 3 gpurun_out/r4z_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r4z_bench_reference.json 2> gpurun_out/r4z_bench_reference.err; head -c 400 gpurun_out/r4z_bench_reference.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r4z_launches_pong.csv python bench.py --profile-step > gpurun_out/r4z_launches.log 2>&1; tail -n 2 gpurun_out/r4z_launches.log
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:tc3 -c 44 -o gpurun_out/r4z_tc3 -f python bench.py --profile-step > gpurun_out/r4z_ncu_tc3.log 2>&1; tail -n 2 gpurun_out/r4z_ncu_tc3.log
ncu -i gpurun_out/r4z_tc3.ncu-rep --page raw --csv > gpurun_out/r4z_tc3_raw.csv 2>/dev/null
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r4z_hbm_pong_learn.csv python bench.py --profile-step > /dev/null 2>&1
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r4z_hbm_navlaser_learn.csv python bench.py --profile-step --workload navlaser > /dev/null 2>&1
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r4z_hbm_pong_fwd.csv python bench.py --profile-forward --steps 1 --no-cpu --no-others > gpurun_out/r4z_hbm_fwd.log 2>&1; tail -n 3 gpurun_out/r4z_hbm_fwd.log
DDRL_LIB_PATH=ddrl4nav_b200/libddrl_b200_timing_ps.so timeout 300 python scratch/tc3_roles_ps.py > gpurun_out/r4z_roles_ps.txt 2>&1; cat gpurun_out/r4z_roles_ps.txt
for f in gpurun_out/*.ncu-rep; do [ $(stat -c %s $f) -gt 40000000 ] && rm -f $f; done
du -sh gpurun_out
