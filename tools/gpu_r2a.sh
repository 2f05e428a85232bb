#!/bin/bash
# round 2, run A: GPU test suite + first c2 bench line
mkdir -p gpurun_out
(nproc; free -g; nvidia-smi -L) > gpurun_out/r2a_host.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -25 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench_c2.json
tail -5 gpurun_out/r2a_bench_c2.err
