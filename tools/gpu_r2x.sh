#!/bin/bash
mkdir -p gpurun_out
for v in under stream under stream; do
CUBEP3M_B200_HISTZERO=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-profile 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), round(d['device_ms_per_step'],3), d['stage_ms_last_step']['link'], d['stage_ms_last_step']['pp_ext'])"
done
