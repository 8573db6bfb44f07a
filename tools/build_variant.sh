#!/bin/bash
# tools/build_variant.sh NAME "-DFOO=1 ..." : experiment build of libcpf.so into build/NAME/ (select with CPF_LIB=build/NAME/libcpf.so)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
OUT=$ROOT/build/$NAME
mkdir -p $OUT/csrc $OUT/include
cp $ROOT/cudaparticlesfoam_b200/csrc/*.cu $ROOT/cudaparticlesfoam_b200/csrc/*.cuh $ROOT/cudaparticlesfoam_b200/csrc/*.h $OUT/csrc/
cp $ROOT/include/cpf.h $OUT/include/
sed -i 's#"../../include/cpf.h"#"../include/cpf.h"#' $OUT/csrc/cpf_internal.h
cd $OUT/csrc
for f in cpf_api cpf_mesh cpf_locate cpf_advect cpf_sort cpf_output cpf_comm; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O2 "$@" -c $f.cu -o $f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libcpf.so *.o -lcudart_static -lpthread -ldl -lrt
echo built $OUT/libcpf.so
