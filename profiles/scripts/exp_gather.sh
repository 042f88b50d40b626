#!/bin/bash
# On the GPU box: the HBM gather (profiles/hbm_gather.py, 64 and 32 wave samples per frame) for the
# default build and every variant under audiality2_b200/build/variants/.
tag=${1:-gather}
mkdir -p gpurun_out
shopt -s nullglob
for lib in default audiality2_b200/build/variants/liba2cu_*.so; do
  n=$(basename $lib .so); n=${n#liba2cu_}
  if [ $lib = default ]; then unset A2CU_LIB; else export A2CU_LIB=$PWD/$lib; fi
  for s in 64 32; do
    timeout 200 python profiles/hbm_gather.py ${VOICES:-131072} 6 $s 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$n', $s, 'samples/frame: %.3f ms  %.1f G  %.0f GB/s  frac %.3f' % (d['ms_per_window'], d['voice_samples_per_s']/1e9, d['algorithmic_GBps'], d['frac']))" | tee -a gpurun_out/${tag}_summary.txt
  done
done
