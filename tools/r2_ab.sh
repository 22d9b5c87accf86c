#!/bin/bash
# round 2 A/B: parity tests + bench + instruction / i-cache counters of one solve launch
# usage: tools/r2_ab.sh <tag> [lib ...]   (default lib = the in-tree libttmpc.so)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
libs=("$@"); [ ${#libs[@]} -eq 0 ] && libs=(trajtrack_mpcndqn_rlboost_b200/libttmpc.so)
MET=smsp__inst_executed.sum,sm__icc_request_hit_rate.pct,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active
for L in "${libs[@]}"; do
  name=$(basename $L .so)
  echo "== $tag $name"
  [ -z "${AB_FAST:-}" ] && TTMPC_LIB=$L python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
  TTMPC_LIB=$L python tools/stress_parity.py 4096 2>&1 | tail -1
  TTMPC_LIB=$L python bench.py --no-cpu-baseline --steps 24 2>/dev/null > gpurun_out/${tag}_${name}_bench.json
  python - gpurun_out/${tag}_${name}_bench.json <<'PY'
import json,sys
l=json.loads([x for x in open(sys.argv[1]) if x.startswith('{')][-1])
print('  in flight %.0f solves/s (%.2f ms/step), e2e %.0f, one batch alone %.1f ms, roofline %.2f %%' % (l['value'], l['ms_per_step'], l['e2e']['value'], l['sequential']['ms_per_step'], 100*l['roofline']['frac']))
PY
  TTMPC_LIB=$L timeout 300 ncu --metrics $MET --clock-control none -k regex:solve_kernel -s 1 -c 1 --csv --log-file gpurun_out/${tag}_${name}_ncu.csv python tools/profile_run.py static4096 2 > /dev/null 2>&1
  python - gpurun_out/${tag}_${name}_ncu.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print('   ', d.get('Metric Name'), d.get('Metric Value'))
PY
done
