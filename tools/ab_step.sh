#!/bin/bash
# A/B of the step-level optimisations on ONE GPU (each variant in its own process: the switches are read once).
# Usage: tools/ab_step.sh  -> gpurun_out/ab_step.txt
mkdir -p gpurun_out
OUT=gpurun_out/ab_step.txt
: > $OUT
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" timeout 200 python bench.py --steps 200 --warmup 20 "$@" 2>gpurun_out/ab_err.log | tail -1)
  echo "$name $(echo "$line" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("ms_per_step=%.4f img_s=%.0f e2e=%.0f b2b=%.4f launches=%s buckets=%s sm=%s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["back_to_back_ms_per_step"], d["launches_per_step"], d["config"].get("buckets"), d["clocks"]["sm_mhz"]))
except Exception as e:
    print("FAILED", e)')" | tee -a $OUT
}
run base            HZ_OVERLAP_ADAM=0 HZ_FUSE_RESADD=0 HZ_PREFETCH_B=0 --
run prefetch        HZ_OVERLAP_ADAM=0 HZ_FUSE_RESADD=0 HZ_PREFETCH_B=1 --
run pf+resadd       HZ_OVERLAP_ADAM=0 HZ_FUSE_RESADD=1 HZ_PREFETCH_B=1 --
run pf+ra+adam296   HZ_OVERLAP_ADAM=1 HZ_FUSE_RESADD=1 HZ_PREFETCH_B=1 --
run all_b4_296      HZ_OVERLAP_ADAM=1 HZ_FUSE_RESADD=1 HZ_PREFETCH_B=1 -- --bucket_mb 4
run all_b4_148      HZ_OVERLAP_ADAM=1 HZ_FUSE_RESADD=1 HZ_PREFETCH_B=1 HZ_ADAM_BUCKET_CTAS=148 -- --bucket_mb 4
run all_b2_148      HZ_OVERLAP_ADAM=1 HZ_FUSE_RESADD=1 HZ_PREFETCH_B=1 HZ_ADAM_BUCKET_CTAS=148 -- --bucket_mb 2
run all_b4_74       HZ_OVERLAP_ADAM=1 HZ_FUSE_RESADD=1 HZ_PREFETCH_B=1 HZ_ADAM_BUCKET_CTAS=74 -- --bucket_mb 4
