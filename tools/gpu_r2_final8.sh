#!/bin/bash
# round-2 final evidence on 8 GPUs: multi-rank parity tests, then the default bench line (configs[3]: 8 x 512^3 particles, 2048^3 mesh)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/r2f_pytest_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_8gpu.log
tail -4 gpurun_out/r2f_pytest_8gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2f_bench_c2_8gpu.json 2> gpurun_out/r2f_bench_c2_8gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_c2_8gpu.json'))
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d.get('cic_power') and d['cic_power'].get('ms'))
print(d['stage_ms_last_step'])
PY
