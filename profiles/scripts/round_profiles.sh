#!/bin/bash
# On the GPU box: the evidence kept under profiles/ for one build - GPU test log, bench lines (ours and
# the reference arm), launch list, ncu --set full of render_split on cfg2 and of the raw-tap gather at
# 64 and 32 wave samples per frame.   round_profiles.sh TAG    -> gpurun_out/TAG_*
tag=${1:-r02b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${tag}_gpu_tests.txt
python bench.py --steps 20 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-configs --no-cpu-baseline --banks 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_split -s 40 -c 1 -f -o gpurun_out/${tag}_split \
    python bench.py --steps 1 --warmup 1 --no-configs --no-cpu-baseline --banks 4 > /dev/null 2>&1
for s in 64 32; do
  ncu --set full --clock-control none --import-source on -k regex:render_split -s 1 -c 1 -f -o gpurun_out/${tag}_gather$s \
      python profiles/hbm_gather.py 131072 2 $s > gpurun_out/${tag}_gather${s}.json 2>&1
done
tail -c 600 gpurun_out/${tag}_gpu_tests.txt
python - <<PY
import json
for l in open('gpurun_out/${tag}_bench.json'):
    try: d = json.loads(l)
    except Exception: continue
    print(round(d['value'] / 1e9, 2), round(d['e2e']['value'] / 1e9, 2), {k: round(v['value'] / 1e9, 1) for k, v in d['configs'].items()},
          {k: round(v['roofline']['frac'], 3) for k, v in d['configs'].items() if 'roofline' in v})
PY
