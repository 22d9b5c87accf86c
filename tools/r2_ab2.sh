#!/bin/bash
# quick A/B: for each "name:lib[:ENV=val,...]" spec run bit-exactness, bench and one-launch counters
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
MET=smsp__inst_executed.sum,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,idc__requests.sum,idc__request_hit_rate.pct
for spec in "$@"; do
  IFS=: read -r name L envs <<< "$spec"
  [ -z "$L" ] && L=trajtrack_mpcndqn_rlboost_b200/libttmpc.so
  EX="TTMPC_LIB=$L"
  [ -n "${envs:-}" ] && EX="$EX ${envs//,/ }"
  echo "== $tag $name ($EX)"
  env $EX python tools/stress_parity.py 2048 2>&1 | tail -1 | cut -c1-120
  env $EX python bench.py --no-cpu-baseline --steps 24 2>/dev/null > gpurun_out/${tag}_${name}_bench.json
  python - gpurun_out/${tag}_${name}_bench.json <<'PY'
import json,sys
try:
    l=json.loads([x for x in open(sys.argv[1]) if x.startswith('{')][-1])
    print('  in flight %.0f solves/s (%.2f ms/step), e2e %.0f, one batch alone %.1f ms' % (l['value'], l['ms_per_step'], l['e2e']['value'], l['sequential']['ms_per_step']))
except Exception as e: print('  bench failed', e)
PY
  env $EX timeout 300 ncu --metrics $MET --clock-control none -k regex:solve_kernel -s 1 -c 1 --csv --log-file gpurun_out/${tag}_${name}_ncu.csv python tools/profile_run.py static4096 2 > /dev/null 2>&1
  python - gpurun_out/${tag}_${name}_ncu.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]; out=[]
short={'gpu__time_duration.sum':'ms','smsp__inst_executed.sum':'Ginst','sm__icc_request_hit_rate.pct':'icc_hit','sm__icc_requests.sum':'icc_req_M','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio':'no_inst','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio':'wait','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio':'short_sb','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio':'br','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active':'fp64%','smsp__issue_active.avg.pct_of_peak_sustained_active':'issue%','idc__requests.sum':'idc_req_M','idc__request_hit_rate.pct':'idc_hit'}
for r in rows[1:]:
    d=dict(zip(h,r)); n=d.get('Metric Name'); v=float(d.get('Metric Value').replace(',',''))
    if n=='gpu__time_duration.sum': v/=1e6
    if n in('smsp__inst_executed.sum',): v/=1e9
    if n in('sm__icc_requests.sum','idc__requests.sum'): v/=1e6
    out.append('%s %.2f'%(short.get(n,n),v))
print('   ', ' | '.join(out))
PY
done
