#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_power.py -m gpu -x -q -k "slab_and_four_step or 1024_mesh or device_cic_power" --durations=5 > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -30 gpurun_out/r2o_pytest.log
