#!/bin/bash
# multi-GPU visit (gpurun --gpus N): multi-rank parity test against the multi-rank oracle, then the weak-scaling bench line at N GPUs
N=${1:-2}
T=${2:-s5m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/${T}_pytest_${N}gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_${N}gpu.log
tail -n 3 gpurun_out/${T}_pytest_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_c1_${N}gpu.json 2> gpurun_out/${T}_bench_c1_${N}gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_c1_${N}gpu.json"))
print("c1 x$N:", round(d["ms_per_step"],3), "ms/step", "%.3e"%d["value"], d["stage_ms_last_step"])
PY
