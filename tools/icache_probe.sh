#!/bin/bash
# Diagnostic: instruction-cache hit rate / fetch stalls of the solve kernel vs resident blocks per SM.
M=sm__icc_request_hit_rate.pct,sm__icc_requests.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_active.avg
for cfg in "1 8192" "2 8192" "3 8192" "1 148"; do
  set -- $cfg
  echo "== blocks/SM cap $1, scenes $2"
  TTMPC_NO_HELPERS=1 TTMPC_MAX_BLOCKS_PER_SM=$1 timeout 200 ncu --metrics $M --clock-control none -k regex:solve_kernel -s 1 -c 1 python tools/profile_run.py static4096 2 $2 2>&1 | grep -E "icc|no_instruction|stalled_wait|short_score|issue_active|time_duration|inst_executed|cycles_active"
done
