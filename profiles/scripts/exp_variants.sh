#!/bin/bash
# On the GPU box: role profile + short bench (+ the HBM gather) of the default build and of every
# variant under audiality2_b200/build/variants/ (profiles/scripts/build_variant.sh).
#   exp_variants.sh TAG [gather]      output -> gpurun_out/TAG_*.txt
tag=${1:-exp}
mkdir -p gpurun_out
show() { python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,2), 'G value', round(d['details']['ms_per_window']*1e3,2), 'us/window |', round(d['e2e']['value']/1e9,2), 'G e2e', {k: round(v['value']/1e9,1) for k,v in d.get('configs',{}).items()})"; }
shopt -s nullglob
for lib in default audiality2_b200/build/variants/liba2cu_*.so; do
  n=$(basename $lib .so); n=${n#liba2cu_}
  if [ $lib = default ]; then unset A2CU_LIB; else export A2CU_LIB=$PWD/$lib; fi
  echo "== $n" | tee -a gpurun_out/${tag}_summary.txt
  timeout 120 python profiles/role_profile.py 4096 > gpurun_out/${tag}_roles_$n.txt 2>&1
  head -8 gpurun_out/${tag}_roles_$n.txt | tee -a gpurun_out/${tag}_summary.txt
  timeout 300 python bench.py --steps 20 --no-cpu-baseline $BENCH_FLAGS 2>/dev/null | tail -1 | show $n | tee -a gpurun_out/${tag}_summary.txt
  if [ "$2" = gather ]; then
    for s in 64 32; do timeout 200 python profiles/hbm_gather.py 131072 6 $s 2>&1 | tail -1 | cut -c1-420 | tee -a gpurun_out/${tag}_summary.txt; done
  fi
done
