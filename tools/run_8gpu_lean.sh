#!/bin/bash
# Lean 8-GPU run (GPU budget: every second costs 8x): bench (default / DDP-like buckets / NCCL all-reduce), fused-TP check
# + in-kernel timeline, the tensor-parallel and data-parallel trainers at world_size 8, DP equivalence, step timeline,
# hybrid dp2 x tp4 and dp2 x pp4 meshes.
N=8; mkdir -p gpurun_out; R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for tag in default c96 size nccl; do
  export HZ_COMM_BLOCKS=32
  case $tag in default) X="";; c96) X=""; export HZ_COMM_BLOCKS=96;; size) X="--bucket_layout size";; nccl) X="--bucket_layout size --allreduce nccl";; esac
  timeout 90 $R --nproc-per-node $N --master-port 2957$((RANDOM % 10)) bench.py --gpus $N --steps 200 --warmup 20 $X > gpurun_out/bench${N}_$tag.log 2>&1; echo "bench$N $tag: $(grep '"metric"' gpurun_out/bench${N}_$tag.log | tail -1 | cut -c1-230)"
done
unset HZ_COMM_BLOCKS
timeout 150 $R --nproc-per-node $N --master-port 29562 tools/tp_fused_check.py gpurun_out/tp_fused$N.json > gpurun_out/tp_fused$N.log 2>&1; grep -E "^heap|FAIL|tp_fused:" gpurun_out/tp_fused$N.log | head -8
timeout 90 $R --nproc-per-node $N --master-port 29563 tools/tp_timeline.py gpurun_out/tp_timeline$N.json > gpurun_out/tp_timeline$N.log 2>&1; grep "^{" gpurun_out/tp_timeline$N.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['case'], d['variant'], round(d['graph_us_per_call_max_over_ranks'], 2), d.get('phase_us_mean', ''))"
timeout 100 python tensor_parallel_train.py --world_size $N --epochs 3 --sample_size 8192 --logs_dir gpurun_out/logs_tp$N > gpurun_out/tp$N.log 2>&1; grep -E "Epoch \[3|Error|completed in|Traceback" gpurun_out/tp$N.log | head -3
timeout 100 python data_parallel_train.py --world_size $N --epochs 3 --sample_size 50000 --logs_dir gpurun_out/logs_dp$N > gpurun_out/dp$N.log 2>&1; grep -E "Epoch \[3|Error|completed in|Traceback" gpurun_out/dp$N.log | head -4
timeout 100 $R --nproc-per-node $N --master-port 29564 tools/equiv_check.py dp gpurun_out/equiv_dp$N.json 2>&1 | grep EQUIV | cut -c1-600
timeout 90 $R --nproc-per-node $N --master-port 29566 tools/trace_step_dist.py gpurun_out/trace${N}_default 2>&1 | grep "^{" | cut -c1-700
timeout 100 python hybrid_parallel_train.py --world_size $N --dp_replicas 2 --inner tensor --epochs 2 --sample_size 16384 --logs_dir gpurun_out/logs_hybrid_dp2tp4 > gpurun_out/hybrid_tp$N.log 2>&1; grep -E "Epoch \[2|Error|completed in|Traceback" gpurun_out/hybrid_tp$N.log | head -3
timeout 100 python hybrid_parallel_train.py --world_size $N --dp_replicas 2 --inner layer --epochs 2 --sample_size 16384 --logs_dir gpurun_out/logs_hybrid_dp2pp4 > gpurun_out/hybrid_pp$N.log 2>&1; grep -E "Epoch \[2|Error|completed in|Traceback" gpurun_out/hybrid_pp$N.log | head -3
python - <<'PY'
import pandas as pd, glob, json
for d in ("logs_tp8", "logs_dp8", "logs_hybrid_dp2tp4", "logs_hybrid_dp2pp4"):
    for f in glob.glob(f"gpurun_out/{d}/combined_results_*.csv"):
        df = pd.read_csv(f)
        cols = [c for c in ["loss", "accuracy", "images_per_sec", "compute_time", "comm_time", "fwd_ms", "bwd_ms", "allreduce_ms", "p2p_ms", "exposed_comm_ms"] if c in df.columns]
        print(d); print(df.groupby("epoch")[cols].max().tail(1).to_string())
    for f in glob.glob(f"gpurun_out/{d}/summary_*.json"):
        s = json.load(open(f)); print("   ", {k: v for k, v in s.items() if k in ("fused_tp", "library_collectives_in_step", "native_fallbacks", "graph", "graph_error", "bucket_algos", "mesh")})
PY
