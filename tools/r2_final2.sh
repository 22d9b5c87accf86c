#!/bin/bash
# Third session of round 2: everything BENCH.md / profiles/ quote for the final build, one GPU, one call.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
bash tools/r2_final.sh
bash tools/r2_single_sweep.sh
# steady state: the same protocol with 100 steps (the drain at the end of the timed region weighs 5x less)
python bench.py --no-cpu-baseline --quick --steps 100 --warmup 5 2>/dev/null > gpurun_out/r2_bench_static4096_100steps.json; python tools/bench_brief.py "static4096 100 steps" < gpurun_out/r2_bench_static4096_100steps.json
python bench.py --no-cpu-baseline --quick --workload mixed4096 --steps 100 --warmup 5 2>/dev/null > gpurun_out/r2_bench_mixed4096_100steps.json; python tools/bench_brief.py "mixed4096 100 steps" < gpurun_out/r2_bench_mixed4096_100steps.json
