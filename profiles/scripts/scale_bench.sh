#!/bin/bash
# On an N-GPU box (gpurun --gpus N): bench.py at 1, 2, 4 ... N GPUs, launched as the driver does.
#   scale_bench.sh TAG N     -> gpurun_out/TAG_n{1,2,4,8}_bench.json
tag=${1:-scale}; nmax=${2:-8}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --no-cpu-baseline > gpurun_out/${tag}_n1_bench.json 2> gpurun_out/${tag}_n1_bench.err
for n in 2 4 8; do
  [ $n -le $nmax ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 > gpurun_out/${tag}_n${n}_bench.json 2> gpurun_out/${tag}_n${n}_bench.err
done
python - <<PY
import json, glob
base = None
for n in (1, 2, 4, 8):
    try:
        d = [json.loads(l) for l in open('gpurun_out/${tag}_n%d_bench.json' % n) if l.startswith('{')][-1]
    except Exception as ex:
        continue
    if n == 1: base = (d['value'], d['e2e']['value'])
    print(n, round(d['value'] / 1e9, 1), round(d['e2e']['value'] / 1e9, 1),
          'eff %.3f / %.3f' % (d['value'] / n / base[0], d['e2e']['value'] / n / base[1]) if base else '',
          d['details'].get('parity_check'), {k: round(v['value'] / 1e9, 1) for k, v in d.get('configs', {}).items()})
PY
