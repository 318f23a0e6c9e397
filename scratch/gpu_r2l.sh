#!/bin/bash
# evidence session: ncu launch list, ncu full captures of the tc3 kernels, HBM-kernel DRAM rows, compute-sanitizer
set -x
mkdir -p gpurun_out
nvidia-smi -L
# 1. launch list of the bench command (kernel share of the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-others > gpurun_out/r2l_launches.log 2>&1; tail -2 gpurun_out/r2l_launches.log; wc -l gpurun_out/r2l_launches.csv
# 2. DRAM bytes + duration of the HBM-bound kernels (one learn step per workload; GAE / easybytes through the kernel tests)
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:'ppo_loss|clip_adam|sumsq|colsum|amax|split_f16|pack|unpack|skinny|s2d|finish' -c 400 --csv --log-file gpurun_out/r2l_hbm_pong.csv python scratch/shape_prof.py pong > gpurun_out/r2l_hbm_pong.log 2>&1; wc -l gpurun_out/r2l_hbm_pong.csv
timeout 600 ncu --metrics $M --clock-control none -k regex:'ppo_loss|pool|thin|im2col|col2im|act_bwd|gaussian|categorical' -c 300 --csv --log-file gpurun_out/r2l_hbm_navlaser.csv python scratch/shape_prof.py navlaser > gpurun_out/r2l_hbm_navlaser.log 2>&1; wc -l gpurun_out/r2l_hbm_navlaser.csv
timeout 600 ncu --metrics $M --clock-control none -k regex:'gae|easybytes|head|sample|encode' -c 200 --csv --log-file gpurun_out/r2l_hbm_misc.csv python -m pytest tests/test_gpu_kernels.py tests/test_easybytes.py -x -q -m gpu -k "full_size or easybytes or head or device" > gpurun_out/r2l_hbm_misc.log 2>&1; wc -l gpurun_out/r2l_hbm_misc.csv
# 3. ncu full capture: tensor-pipe activity of every distinct tc3 launch of one Pong learn iteration (raw csv only; reps are too big)
timeout 900 ncu --set full --clock-control none -k regex:'tc3_' -s 60 -c 26 --csv --page raw --log-file gpurun_out/r2l_ncu_tc3_raw.csv python scratch/shape_prof.py pong > gpurun_out/r2l_ncu_tc3.log 2>&1; wc -l gpurun_out/r2l_ncu_tc3_raw.csv
# 4. compute-sanitizer over the kernel tests (small shapes)
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "not full_size" > gpurun_out/r2l_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -n 6 gpurun_out/r2l_sanitizer_$tool.log
done
