#!/bin/bash
mkdir -p gpurun_out
( echo "# round 2 final: compute-sanitizer --tool memcheck python tools/sanitize_run.py"; timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2>&1 | grep -v "^$" | tail -60 ) > gpurun_out/r2_final_compute_sanitizer_memcheck.txt
( echo "# round 2 final: compute-sanitizer --tool racecheck python tools/sanitize_run.py"; timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_run.py 2>&1 | grep -v "^$" | tail -60 ) > gpurun_out/r2_final_compute_sanitizer_racecheck.txt
tail -n 14 gpurun_out/r2_final_compute_sanitizer_memcheck.txt; tail -n 4 gpurun_out/r2_final_compute_sanitizer_racecheck.txt
