#!/bin/bash
# Lean A/B (third session of round 2): per "name:lib" spec one bit-exactness check against the oracle (1024 mixed
# scenes) and ONE run of the driver's bench protocol; one-launch ncu counters only for the specs named in NCU_FOR.
# MIXED=0 skips the mixed4096 run.
# usage: NCU_FOR="cur sr3" tools/r2_ab5.sh <tag> spec...     (lib empty = the product library)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
MET=smsp__inst_executed.sum,sm__icc_request_hit_rate.pct,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
for spec in "$@"; do
  IFS=: read -r name L <<< "$spec"
  [ -z "$L" ] && L=trajtrack_mpcndqn_rlboost_b200/libttmpc.so
  echo "== $tag $name ($L)" | tee -a gpurun_out/${tag}_brief.txt
  TTMPC_LIB=$L python tools/stress_parity.py 1024 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/${tag}_brief.txt
  TTMPC_LIB=$L python bench.py --no-cpu-baseline --quick --steps 20 --warmup 5 2>/dev/null | python tools/bench_brief.py "  $name static20" | tee -a gpurun_out/${tag}_brief.txt
  [ "${MIXED:-1}" = "1" ] && TTMPC_LIB=$L python bench.py --no-cpu-baseline --quick --workload mixed4096 --steps 20 --warmup 5 2>/dev/null | python tools/bench_brief.py "  $name mixed20" | tee -a gpurun_out/${tag}_brief.txt
  case " ${NCU_FOR:-} " in *" $name "*)
    for wl in static4096 mixed4096; do
    TTMPC_LIB=$L timeout 300 ncu --metrics $MET --clock-control none -k regex:solve_kernel -s 1 -c 1 --csv --log-file gpurun_out/${tag}_${name}_${wl}_ncu.csv python tools/profile_run.py $wl 2 > /dev/null 2>&1
    python - gpurun_out/${tag}_${name}_${wl}_ncu.csv $wl <<'PY' | tee -a gpurun_out/${tag}_brief.txt
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
    print('    ncu %s:'%sys.argv[2], ' | '.join(out))
except Exception as e: print('    ncu failed', e)
PY
    done;;
  esac
done
