#!/bin/bash
# 2 GPUs: multi-rank parity (slab coarse solve + peer-memory pass; fallbacks) and a short weak-scaling bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/r2c_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_2gpu.log
tail -30 gpurun_out/r2c_pytest_2gpu.log
for mode in slab replicated; do
CUBEP3M_B200_COARSE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --workload c1 --no-cpu > gpurun_out/r2c_bench_c1_2gpu_$mode.json 2> gpurun_out/r2c_bench_c1_2gpu_$mode.err; echo "bench $mode rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c_bench_c1_2gpu_$mode.json'))
    print('$mode', d['ms_per_step'], d['value'], d['stage_ms_last_step'])
    for k in ('coarse_fft','coarse_misc','coarse_xchg'):
        if k in d['stages']: print(k, d['stages'][k]['ms_per_step'], d['stages'][k]['launches_per_step'])
except Exception as e:
    print('no json', e); print(open('gpurun_out/r2c_bench_c1_2gpu_$mode.err').read()[-2000:])
PY
done
