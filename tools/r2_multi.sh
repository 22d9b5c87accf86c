#!/bin/bash
# round 2, one 8-GPU box: the 1/2/4/8 curve on static4096, BASELINE configs[2] (dynamic8192 per rank) on 8 GPUs,
# the configs[4] sweep shapes on 1 and 8 GPUs (2 and 4 for two of them).  Lines -> gpurun_out/r2_multi_*.json
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # run <n> <tag> <bench args...>
  local n=$1 tag=$2; shift 2
  local port=$((29500 + RANDOM % 1000))
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@" > gpurun_out/r2_multi_${tag}_n$n.json 2> gpurun_out/r2_multi_${tag}_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/r2_multi_${tag}_n$n.json 2> gpurun_out/r2_multi_${tag}_n$n.err; fi
  python - gpurun_out/r2_multi_${tag}_n$n.json $tag $n <<'PY'
import json,sys
try:
    l=json.loads([x for x in open(sys.argv[1]) if x.startswith('{')][-1])
    pr=l.get('per_rank',{})
    print('%s N=%s: %.0f solves/s, %.2f ms/step, e2e %.0f, per-rank ms %s, drain %s' % (sys.argv[2], sys.argv[3], l['value'], l['ms_per_step'], l['e2e']['value'], pr.get('ms_per_step_min_mean_max'), pr.get('drain_ms')))
except Exception as e: print(sys.argv[2], sys.argv[3], 'failed', e)
PY
}
for n in 1 2 4 8; do run $n static4096 --steps 20 --warmup 5 --no-cpu-baseline; done
run 8 dynamic8192 --workload dynamic8192 --steps 6 --warmup 3 --batches 2 --no-cpu-baseline --quick
run 1 dynamic8192 --workload dynamic8192 --steps 6 --warmup 3 --batches 2 --no-cpu-baseline --quick
run 8 mixed4096 --workload mixed4096 --steps 20 --warmup 3 --no-cpu-baseline --quick
for shape in "N=10,Nstc=10,Ndyn=15" "N=32,Nstc=10,Ndyn=15" "N=20,Nstc=4,Ndyn=4" "N=20,Nstc=20,Ndyn=30"; do
  tag=sweep_$(echo $shape | tr ',=' '__')
  for n in ${SWEEP_NS:-1 8}; do run $n $tag --workload "sweep:$shape" --steps 12 --warmup 3 --no-cpu-baseline --quick; done
done
for shape in "N=10,Nstc=10,Ndyn=15" "N=20,Nstc=20,Ndyn=30"; do
  tag=sweep_$(echo $shape | tr ',=' '__')
  for n in 2 4; do run $n $tag --workload "sweep:$shape" --steps 12 --warmup 3 --no-cpu-baseline --quick; done
done
