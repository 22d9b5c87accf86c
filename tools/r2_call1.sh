#!/bin/bash
# round 2, call 1: baseline of the round-1 build -- GPU tests + ncu captures of the workloads that had none
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/r2c1_pytest.log
NCU="ncu --set full --import-source on --clock-control none -k regex:solve_kernel -s 1 -c 1"
timeout 600 $NCU -o gpurun_out/r2_base_mixed4096 -f python tools/profile_run.py mixed4096 2 > gpurun_out/r2c1_ncu_mixed.log 2>&1; echo "ncu mixed rc $?"
timeout 900 $NCU -o gpurun_out/r2_base_dynamic2048 -f python tools/profile_run.py dynamic8192 2 2048 > gpurun_out/r2c1_ncu_dyn.log 2>&1; echo "ncu dyn rc $?"
python bench.py --steps 20 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc $?"
tail -c 600 gpurun_out/r2c1_bench.json
