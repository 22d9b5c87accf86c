#!/bin/bash
# bench lines of several workloads with one library: tools/r2_wl.sh <tag> <lib|-> workload[:steps] ...
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; L=$2; shift 2
[ "$L" = "-" ] && L=trajtrack_mpcndqn_rlboost_b200/libttmpc.so
for spec in "$@"; do
  IFS=: read -r wl steps <<< "$spec"; steps=${steps:-12}
  TTMPC_LIB=$L python bench.py --no-cpu-baseline --workload $wl --steps $steps 2>gpurun_out/${tag}_${wl}.err > gpurun_out/${tag}_${wl}_bench.json
  python - gpurun_out/${tag}_${wl}_bench.json $wl <<'PY'
import json,sys
try:
    l=json.loads([x for x in open(sys.argv[1]) if x.startswith('{')][-1])
    print('%s: in flight %.0f solves/s (%.2f ms/step), e2e %.0f, one batch alone %.1f ms, evals/solve %.0f, status %s' % (sys.argv[2], l['value'], l['ms_per_step'], l['e2e']['value'], l['sequential']['ms_per_step'], l['roofline'].get('evals_per_solve',0), l.get('exit_status_hist')))
except Exception as e: print(sys.argv[2], 'bench failed', e)
PY
done
