#!/bin/bash
# one GPU-box visit: parity tests, the bench line, the reference arm, the ncu launch list and one full capture of the hot kernel
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_rc=$?; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo ref_rc=$?
CMD="python bench.py --steps 1 --warmup 1 --spp 64 --e2e-steps 0 --cpu-spp 0"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo launches_rc=$?
ncu --set full --clock-control none --import-source on -k regex:persist_kernel -s 1 -c 1 -f -o gpurun_out/prof_persist_$TAG $CMD > gpurun_out/ncu_full_persist_$TAG.log 2>&1; echo persist_rc=$?
# DRAM / L2 bytes per path of the captured launch (800 x 800 x 64 paths) -> the file bench.py reads roofline.traffic from
python tools/ncu_traffic.py gpurun_out/prof_persist_$TAG.ncu-rep 40960000 gpurun_out/${TAG}_persist_traffic.json
cat gpurun_out/bench_$TAG.json
