#!/bin/bash
# A/B of the gradient-bucket layout / bucket-wise Adam on N GPUs (default 2).  -> gpurun_out/ab_dpN.txt
N=${1:-2}
mkdir -p gpurun_out
OUT=gpurun_out/ab_dp$N.txt
: > $OUT
PORT=29610
run() {
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  PORT=$((PORT+1))
  line=$(env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 200 --warmup 20 "$@" 2>gpurun_out/ab_dp_err.log | grep '^{"metric' | tail -1)
  echo "$name $(echo "$line" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("ms_per_step=%.4f img_s=%.0f e2e=%.0f b2b=%.4f launches=%s buckets=%s loss=%s sm=%s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["back_to_back_ms_per_step"], d["launches_per_step"], d["config"].get("buckets"), d.get("last_step_loss"), d["clocks"]["sm_mhz"]))
except Exception as e:
    print("FAILED", e)')" | tee -a $OUT
}
run raw25        HZ_BUCKET_LIVE=0 --
run live4        HZ_BUCKET_LIVE=1 -- --live_bucket_mb 4
run live2        HZ_BUCKET_LIVE=1 -- --live_bucket_mb 2
run live1        HZ_BUCKET_LIVE=1 -- --live_bucket_mb 1
run live2_adam   HZ_BUCKET_LIVE=1 HZ_OVERLAP_ADAM=1 -- --live_bucket_mb 2
