#!/bin/bash
# A/B of step-level switches on ONE GPU (each variant in its own process: the switches are read once).
# Usage: tools/ab_step.sh  -> gpurun_out/ab_step.txt
mkdir -p gpurun_out
OUT=gpurun_out/ab_step.txt
: > $OUT
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" timeout 200 python bench.py --steps 200 --warmup 20 "$@" 2>gpurun_out/ab_err_$name.log | tail -1)
  echo "$name $(echo "$line" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("ms_per_step=%.4f img_s=%.0f e2e=%.0f b2b=%.4f launches=%s sm=%s loss=%s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["back_to_back_ms_per_step"], d["launches_per_step"], d["clocks"]["sm_mhz"], d.get("last_step_loss")))
except Exception as e:
    print("FAILED", e)')" | tee -a $OUT
}
run default      --
run mink10       HZ_CLUSTER_MIN_K=10 --
run mink16       HZ_CLUSTER_MIN_K=16 --
run mink4        HZ_CLUSTER_MIN_K=4 --
