#!/bin/bash
# N-GPU sweep of the data-parallel step (bench.py value / ms_per_step per configuration): optimizer placement (separate
# pass, fused into the all-reduce kernel), bucket layout (DDP-like size caps vs layer groups with a small last bucket
# on the flag-in-data latency protocol), CTA caps of the all-reduces that run in the shadow of backward.  NGPU=2|4|8
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node ${NGPU:-2}"
run() { tag=$1; shift; env "$@" timeout 120 $R --master-port $((29600 + RANDOM % 300)) bench.py --gpus ${NGPU:-2} --steps 100 --warmup 10 $EXTRA > gpurun_out/sweep${NGPU:-2}_$tag.log 2>&1; python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/sweep${NGPU:-2}_$tag.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("$tag", d["ms_per_step"], d["value"], d["config"].get("bucket_allreduce_algos"), d["back_to_back_ms_per_step"])
except Exception as e: print("$tag FAILED", e)
PY
}
EXTRA="--bucket_layout layers"; run layers_c32 HZ_COMM_BLOCKS=32; run layers_c96 HZ_COMM_BLOCKS=96
EXTRA="--bucket_layout size"; run size_c96 HZ_COMM_BLOCKS=96
EXTRA="--bucket_layout layers --fused_adam"; run fused_e32 HZ_COMM_BLOCKS_EARLY=32
EXTRA="--bucket_layout size --allreduce nccl"; run nccl_allreduce X=1
