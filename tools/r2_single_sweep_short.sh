#!/bin/bash
# the one-GPU workload rows of BENCH.md except dynamic8192 (tools/r2_single_sweep.sh without its slowest run)
set -u
cd "$(dirname "$0")/.."
source <(sed -n '/^run() {/,/^}/p' tools/r2_multi.sh)
mkdir -p gpurun_out
run 1 mixed4096 --workload mixed4096 --steps 20 --warmup 3 --no-cpu-baseline --quick
for shape in "N=10,Nstc=10,Ndyn=15" "N=20,Nstc=20,Ndyn=30" "N=20,Nstc=4,Ndyn=4" "N=32,Nstc=10,Ndyn=15"; do
  tag=sweep_$(echo $shape | tr ',=' '__')
  run 1 $tag --workload "sweep:$shape" --steps 12 --warmup 3 --no-cpu-baseline --quick
done
