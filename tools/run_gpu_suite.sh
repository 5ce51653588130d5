#!/bin/bash
# Multi-GPU evaluation suite, N = $NGPU (one `gpurun --gpus N` call): collective + fused-TP checks, strategy equivalence,
# bench (default / DDP-like buckets / NCCL all-reduce), the three trainer CLIs at the graded world sizes, hybrid meshes and
# (SWEEP=1) the main.py sweep.  Everything lands in gpurun_out/; summaries are copied into profiles/ afterwards.
N=${NGPU:-2}; PP=$(( N < 4 ? N : 4 )); mkdir -p gpurun_out; R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $R --nproc-per-node $N --master-port 29562 tools/tp_fused_check.py gpurun_out/tp_fused$N.json > gpurun_out/tp_fused$N.log 2>&1; grep -E "^heap|FAIL|tp_fused:" gpurun_out/tp_fused$N.log | head -8
timeout 120 $R --nproc-per-node $N --master-port 29563 tools/tp_timeline.py gpurun_out/tp_timeline$N.json > gpurun_out/tp_timeline$N.log 2>&1; grep "^{" gpurun_out/tp_timeline$N.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['case'], d['variant'], round(d['graph_us_per_call_max_over_ranks'], 2), d.get('phase_us_mean', ''))"
timeout 240 $R --nproc-per-node $N --master-port 29561 tools/multigpu_check.py gpurun_out/multigpu$N.json > gpurun_out/multigpu$N.log 2>&1; grep -E "multigpu:|FAIL" gpurun_out/multigpu$N.log | head -5
timeout 200 $R --nproc-per-node $N --master-port 29564 tools/equiv_check.py dp gpurun_out/equiv_dp$N.json 2>&1 | grep EQUIV | cut -c1-600
timeout 200 $R --nproc-per-node $PP --master-port 29565 tools/equiv_check.py pp gpurun_out/equiv_pp$PP.json 2>&1 | grep EQUIV | cut -c1-600
for tag in default size nccl; do
  case $tag in default) X="";; size) X="--bucket_layout size";; nccl) X="--bucket_layout size --allreduce nccl";; esac
  timeout 120 $R --nproc-per-node $N --master-port 2957$((RANDOM % 10)) bench.py --gpus $N --steps 200 --warmup 20 $X > gpurun_out/bench${N}_$tag.log 2>&1; echo "bench$N $tag: $(grep '"metric"' gpurun_out/bench${N}_$tag.log | tail -1 | cut -c1-230)"
done
timeout 120 python data_parallel_train.py --world_size $N --epochs 3 --sample_size 50000 --logs_dir gpurun_out/logs_dp$N > gpurun_out/dp$N.log 2>&1; grep -E "Epoch \[3|Error|completed in|Traceback" gpurun_out/dp$N.log | head -4
timeout 120 python tensor_parallel_train.py --world_size $N --epochs 3 --sample_size 8192 --logs_dir gpurun_out/logs_tp$N > gpurun_out/tp$N.log 2>&1; grep -E "Epoch \[3|Error|completed in|Traceback" gpurun_out/tp$N.log | head -3
timeout 120 python layer_model_parallel_train.py --world_size $PP --epochs 3 --sample_size 8192 --logs_dir gpurun_out/logs_pp$PP > gpurun_out/pp$PP.log 2>&1; grep -E "Epoch \[3|Error|completed in|Traceback|capture failed" gpurun_out/pp$PP.log | head -4
timeout 120 python layer_model_parallel_train.py --world_size $PP --epochs 3 --sample_size 32768 --batch_size 512 --microbatches 8 --logs_dir gpurun_out/logs_pp${PP}_b512 > gpurun_out/pp${PP}_b512.log 2>&1; grep -E "Epoch \[3|Error|Traceback|capture failed" gpurun_out/pp${PP}_b512.log | head -3
[ "${LIGHT:-0}" = "1" ] || timeout 100 python layer_model_parallel_train.py --world_size $PP --epochs 3 --sample_size 32768 --batch_size 512 --microbatches 8 --no_pp_overlap --logs_dir gpurun_out/logs_pp${PP}_b512_blocking > gpurun_out/pp${PP}_b512_blocking.log 2>&1; grep -E "Epoch \[3|Error|Traceback" gpurun_out/pp${PP}_b512_blocking.log | head -3
[ "${LIGHT:-0}" = "1" ] || timeout 100 python data_parallel_train.py --world_size 1 --epochs 3 --sample_size 32768 --batch_size 512 --logs_dir gpurun_out/logs_dp1_b512 > gpurun_out/dp1_b512.log 2>&1; grep -E "Epoch \[3|Error|Traceback" gpurun_out/dp1_b512.log | head -3
if [ $N -ge 4 ]; then
  timeout 100 python hybrid_parallel_train.py --world_size $N --dp_replicas 2 --inner layer --epochs 2 --sample_size 16384 --logs_dir gpurun_out/logs_hybrid_dp2pp$((N/2)) > gpurun_out/hybrid_pp$N.log 2>&1; grep -E "Epoch \[2|Error|completed in|Traceback" gpurun_out/hybrid_pp$N.log | head -3
  timeout 100 python hybrid_parallel_train.py --world_size $N --dp_replicas 2 --inner tensor --epochs 2 --sample_size 16384 --logs_dir gpurun_out/logs_hybrid_dp2tp$((N/2)) > gpurun_out/hybrid_tp$N.log 2>&1; grep -E "Epoch \[2|Error|completed in|Traceback" gpurun_out/hybrid_tp$N.log | head -3
fi
if [ $N -ge 8 ]; then
  timeout 100 $R --nproc-per-node $N --master-port 29566 tools/trace_step_dist.py gpurun_out/trace${N}_default 2>&1 | grep "^{" | cut -c1-600
  timeout 100 $R --nproc-per-node $N --master-port 29567 tools/trace_step_dist.py gpurun_out/trace${N}_size --size 2>&1 | grep "^{" | cut -c1-600
fi
if [ "${SWEEP:-0}" = "1" ]; then
  timeout 900 python main.py --sample_sizes 1000 10000 50000 --world_size $N --epochs 2 --output_dir gpurun_out/benchmark_results_${N}gpu > gpurun_out/main$N.log 2>&1; tail -3 gpurun_out/main$N.log
fi
python - <<'PY'
import pandas as pd, glob, json
for d in sorted(glob.glob("gpurun_out/logs_*")):
    for f in glob.glob(f"{d}/combined_results_*.csv"):
        df = pd.read_csv(f)
        cols = [c for c in ["loss", "accuracy", "images_per_sec", "compute_time", "comm_time", "fwd_ms", "bwd_ms", "allreduce_ms", "p2p_ms", "exposed_comm_ms"] if c in df.columns]
        print(d.split("/")[-1]); print(df.groupby("epoch")[cols].max().tail(1).to_string())
    for f in glob.glob(f"{d}/summary_*.json"):
        s = json.load(open(f)); print("   ", {k: v for k, v in s.items() if k in ("fused_tp", "library_collectives_in_step", "native_fallbacks", "graph", "graph_error", "fused_adam", "bucket_algos", "mesh")})
PY
