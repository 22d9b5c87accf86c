#!/bin/bash
# Round-2 A/B of the opt-in builds against the default library, one gpurun call:
#   tools/build_variant.sh cold -DTTMPC_COLD_OUTLINE        (here, on the CPU box; build/ travels)
#   gpurun -- 'bash tools/variants_r2.sh cold'
# For every variant: bit-exactness (stress_parity on 4096 mixed scenes through the bulk kernel),
# bench.py with six batches in flight, single-scene latency; the default library first.
set -u
cd "$(dirname "$0")/.."
for v in default "$@"; do
  L=trajtrack_mpcndqn_rlboost_b200/libttmpc.so
  [ "$v" != default ] && L=build/libttmpc_$v.so
  echo "== $v ($L)"
  TTMPC_LIB=$L python tools/stress_parity.py 4096 2>&1 | tail -1
  TTMPC_LIB=$L python bench.py --no-cpu-baseline --steps 24 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin if x.startswith('{')][-1])
print('  in flight %.0f solves/s (%.2f ms/step), e2e %.0f, one batch alone %.1f ms, roofline %.2f %%' % (l['value'], l['ms_per_step'], l['e2e']['value'], l['sequential']['ms_per_step'], 100*l['roofline']['frac']))"
  TTMPC_LIB=$L python tools/single_latency.py 2>&1 | tail -2 | head -1
done
