#!/bin/bash
# A/B of bench.py with an engine switch (env var) on the GPU box:  ab_bench.sh A2CU_NO_COPY_STREAM
show() { python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,2), 'G value', round(d['ms_per_step']*1e3,1), 'us |', round(d['e2e']['value']/1e9,2), 'G e2e', round(d['e2e']['ms_per_step']*1e3,1), 'us')"; }
for i in 1 2; do timeout 100 python bench.py --steps 400 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | show on; done
env $1=1 timeout 100 python bench.py --steps 400 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | show "off($1)"
