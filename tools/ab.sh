#!/bin/bash
# tools/ab.sh TAG variant... : bench.py (20+5 steps, no CPU leg) for the in-tree library ("base") and experiment builds
TAG=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset CPF_LIB; else export CPF_LIB=build/$v/libcpf.so; fi
  python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err
  python tools/pr.py gpurun_out/${TAG}_$v.json
done
