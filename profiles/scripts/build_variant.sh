#!/bin/bash
# Build a variant of liba2cu.so with extra -D flags on the render_split translation unit:
#   profiles/scripts/build_variant.sh pf32 -DA2CU_SER_PF=32
# -> audiality2_b200/build/variants/liba2cu_pf32.so (travels with the gpurun snapshot, git-ignored).
# A script selects it with  A2CU_LIB=<path>  (audiality2_b200/engine.py).
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
d=audiality2_b200/build
mkdir -p $d/variants
python -m audiality2_b200.build > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-extended-lambda -Xcompiler -fPIC \
     -cudart shared -Xfatbin -compress-all -I include "$@" -c -o $d/variants/split_$name.o audiality2_b200/csrc/a2cu_reg_split.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart shared -o $d/variants/liba2cu_$name.so \
     $d/a2cu_engine.o $d/a2cu_reg_bank_wt.o $d/a2cu_reg_bank_fm.o $d/variants/split_$name.o
echo $d/variants/liba2cu_$name.so
