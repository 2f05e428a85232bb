#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_init.py tests/test_checkpoint.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -30 gpurun_out/r2g_pytest.log
