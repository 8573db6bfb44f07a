timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extensions.py tests/test_gpu_configs.py tests/test_gpu_honeycomb.py -x -q -m gpu 2>&1 | tail -3
run() { tag=$1; shift; env "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err; python tools/pr.py gpurun_out/$tag.json | cut -c1-330; }
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-extra"
run l_c2 X=1 $B --workload box100_1e6
run l_esc X=1 $B --workload channel1M_1e7_escape
run l_head X=1 $B
