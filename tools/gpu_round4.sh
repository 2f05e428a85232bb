#!/bin/bash
# GPU visit: full parity suite (incl. the newest PP_EXT tests) and smoke()
mkdir -p gpurun_out
T=${1:-s5h}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -n 15 gpurun_out/${T}_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -n 2 gpurun_out/${T}_smoke.log
