#!/bin/bash
# Lean A/B on the GPU box: for each "name:lib[:ENV=val,...]" spec -- bit-exactness against the oracle (1024 mixed
# scenes), the driver's bench protocol twice (--steps 20 --warmup 5, quick), one-launch ncu counters.
# usage: tools/r2_ab3.sh <tag> spec...        (lib empty = the product library; optional 4th field: extra bench.py arguments, comma separated)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
MET=smsp__inst_executed.sum,sm__icc_request_hit_rate.pct,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
for spec in "$@"; do
  IFS=: read -r name L envs bargs <<< "$spec"
  [ -z "$L" ] && L=trajtrack_mpcndqn_rlboost_b200/libttmpc.so
  EX="TTMPC_LIB=$L"
  [ -n "${envs:-}" ] && EX="$EX ${envs//,/ }"
  echo "== $tag $name ($EX)"
  env $EX python tools/stress_parity.py 1024 2>&1 | tail -1 | cut -c1-150
  for rep in 1 2; do
    env $EX python bench.py --no-cpu-baseline --quick --steps 20 --warmup 5 ${bargs//,/ } 2>/dev/null | python tools/bench_brief.py "  $name#$rep" | tee -a gpurun_out/${tag}_brief.txt
  done
  if [ "${AB_NCU:-1}" = "1" ]; then
    env $EX timeout 300 ncu --metrics $MET --clock-control none -k regex:solve_kernel -s 1 -c 1 --csv --log-file gpurun_out/${tag}_${name}_ncu.csv python tools/profile_run.py static4096 2 > /dev/null 2>&1
    python - gpurun_out/${tag}_${name}_ncu.csv <<'PY' | tee -a gpurun_out/${tag}_brief.txt
import csv,sys
try:
    rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
    h=rows[0]; out=[]
    short={'gpu__time_duration.sum':'ms','smsp__inst_executed.sum':'Ginst','sm__icc_request_hit_rate.pct':'icc_hit','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio':'no_inst','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio':'wait','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio':'short_sb','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active':'fp64%','smsp__issue_active.avg.pct_of_peak_sustained_active':'issue%'}
    for r in rows[1:]:
        d=dict(zip(h,r)); n=d.get('Metric Name'); v=float(d.get('Metric Value').replace(',',''))
        if n=='gpu__time_duration.sum': v/=1e6
        if n=='smsp__inst_executed.sum': v/=1e9
        out.append('%s %.2f'%(short.get(n,n),v))
    print('    ncu:', ' | '.join(out))
except Exception as e: print('    ncu failed', e)
PY
  fi
done
