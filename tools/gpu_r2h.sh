#!/bin/bash
# 8 GPUs: configs[3]/[4] shape (512^3 particles per GPU, nodes_dim = 2), multi-rank parity test, short bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k native > gpurun_out/r2h_pytest_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest_8gpu.log
tail -5 gpurun_out/r2h_pytest_8gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu > gpurun_out/r2h_bench_c2_8gpu.json 2> gpurun_out/r2h_bench_c2_8gpu.err; echo "bench rc=$?"
tail -5 gpurun_out/r2h_bench_c2_8gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench_c2_8gpu.json'))
print(d['ms_per_step'], d['value'], d['device_ms_per_step'], d['e2e'], d['stage_ms_last_step'])
for k in ('coarse_fft','coarse_misc','coarse_xchg','pass_pack','pass_unpack'):
    if k in d['stages']: print(k, d['stages'][k]['ms_per_step'], d['stages'][k]['launches_per_step'])
PY
nvidia-smi topo -m > gpurun_out/r2h_topo.txt 2>&1
