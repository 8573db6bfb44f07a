#!/bin/bash
# tools/gpu_profile.sh TAG : launch list + one ncu --set full capture of the all-particles kernel (k_lean) of the bench command
TAG=${1:-prof}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lean --launch-skip 3 -c 1 -f -o gpurun_out/${TAG}_k_lean \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out/
