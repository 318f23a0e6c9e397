#!/bin/bash
# One dense GPU call: parity tests, three bench lines, ncu launch list, ncu full capture of the tc2 engine and GAE.
# gpurun only copies gpurun_out/ back when it is <= 64 MiB: raw CSV pages are exported on the box and large reports dropped.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_pong.json 2> gpurun_out/bench_pong.err; tail -c 3000 gpurun_out/bench_pong.json
timeout 600 python bench.py --workload navlaser > gpurun_out/bench_navlaser.json 2> gpurun_out/bench_navlaser.err
timeout 600 python bench.py --workload navimg > gpurun_out/bench_navimg.json 2> gpurun_out/bench_navimg.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_pong.csv python bench.py --profile-step > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc2 -c 14 -o gpurun_out/r1_tc2 -f python bench.py --profile-step > gpurun_out/ncu_tc2.log 2>&1
ncu -i gpurun_out/r1_tc2.ncu-rep --page raw --csv > gpurun_out/r1_tc2_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:gae -s 1 -c 2 -o gpurun_out/r1_gae -f python -c "import torch,sys; sys.path.insert(0,'.'); from ddrl4nav_b200 import kernels; T,N=2048,65536; d='cuda'; v=torch.randn(T+1,1,N,device=d); r=torch.randn(T,1,N,device=d); dn=(torch.rand(T,1,N,device=d)<0.02).to(torch.uint8); [kernels.gae(v,r,dn,[0.99],0.95) for _ in range(3)]; torch.cuda.synchronize()"  > gpurun_out/ncu_gae.log 2>&1
ncu -i gpurun_out/r1_gae.ncu-rep --page raw --csv > gpurun_out/r1_gae_raw.csv 2>/dev/null
timeout 300 python scratch/shape_prof.py pong > gpurun_out/shape_pong.txt 2>&1
timeout 300 python scratch/shape_prof.py navlaser > gpurun_out/shape_navlaser.txt 2>&1
timeout 300 python scratch/shape_prof.py navimg > gpurun_out/shape_navimg.txt 2>&1
# stay under the copy-back limit
for f in gpurun_out/*.ncu-rep; do [ $(stat -c %s $f) -gt 30000000 ] && rm -f $f; done
du -sh gpurun_out; ls -la gpurun_out
