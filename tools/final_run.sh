#!/bin/bash
# tools/final_run.sh TAG : the measurements behind profiles/<TAG>_* on ONE B200 (tests, default bench line, reference arm, variants,
# BASELINE configs 4 and 5 on one GPU, launch list, ncu --set full of the two kernels of the advect launch sequence)
TAG=${1:-final}
O=gpurun_out
mkdir -p $O
python -m pytest tests -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; tail -2 $O/${TAG}_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; python tools/pr.py $O/${TAG}_bench_n1.json | cut -c1-140
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_ref.err; tail -c 600 $O/${TAG}_bench_reference_arm.json
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-extra"
$B --rng xorwow --fuse 0 > $O/${TAG}_var_rngxorwowfuse0.json 2>/dev/null
$B --integrator rk2 > $O/${TAG}_var_integratorrk2.json 2>/dev/null
$B --integrator rk4 > $O/${TAG}_var_integratorrk4.json 2>/dev/null
$B --interp vertex > $O/${TAG}_var_interpvertex.json 2>/dev/null
$B --locator bary > $O/${TAG}_var_locatorbary.json 2>/dev/null
$B --exact > $O/${TAG}_var_exact.json 2>/dev/null
python tools/pr.py $O/${TAG}_var_*.json | cut -c1-120
python bench.py --workload channel1M_1e8_escape --full-single --steps 10 --warmup 3 --no-cpu --no-extra > $O/${TAG}_c5_n1.json 2>/dev/null
python bench.py --workload poly10M_1e8_rk4 --steps 5 --warmup 3 --no-cpu --no-extra > $O/${TAG}_c4_n1.json 2>/dev/null
python tools/pr.py $O/${TAG}_c5_n1.json $O/${TAG}_c4_n1.json | cut -c1-120
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 400 --csv \
    --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lean --launch-skip 3 -c 1 -f -o $O/${TAG}_k_lean python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fast --launch-skip 3 -c 1 -f -o $O/${TAG}_k_fin python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
ls $O | grep ${TAG} | wc -l
