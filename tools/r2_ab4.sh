#!/bin/bash
# Third session of round 2: TT_OPT (scalar work of the PANOC step) against the build without it.
# usage (GPU box): tools/r2_ab4.sh      needs build/libttmpc_opt0.so (tools/build_variant.sh opt0 -DTT_OPT=0)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_NCU=1 tools/r2_ab3.sh ab7 base:build/libttmpc_opt0.so new:
for spec in base:build/libttmpc_opt0.so new:trajtrack_mpcndqn_rlboost_b200/libttmpc.so; do
  IFS=: read -r name L <<< "$spec"
  TTMPC_LIB=$L python bench.py --no-cpu-baseline --quick --workload mixed4096 --steps 20 --warmup 5 2>/dev/null | python tools/bench_brief.py "  mixed20 $name" | tee -a gpurun_out/ab7_brief.txt
done
python bench.py --no-cpu-baseline --quick --workload mixed4096 --steps 100 --warmup 5 2>/dev/null | python tools/bench_brief.py "  mixed100 new" | tee -a gpurun_out/ab7_brief.txt
python bench.py --no-cpu-baseline --quick --steps 100 --warmup 5 2>/dev/null | python tools/bench_brief.py "  static100 new" | tee -a gpurun_out/ab7_brief.txt
