#!/bin/bash
# 2-GPU validation suite (one `gpurun --gpus 2` call): fused TP kernels incl. the NVSwitch-multicast variant, collective
# checks, the three trainers at world_size 2, bench at 2 GPUs (fused all-reduce+Adam vs the separate optimizer pass).
mkdir -p gpurun_out; R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $R --nproc-per-node 2 --master-port 29562 tools/tp_fused_check.py gpurun_out/tp_fused2.json > gpurun_out/tp_fused2.log 2>&1; grep -E "^heap|FAIL|tp_fused:|Error|error" gpurun_out/tp_fused2.log | head -12; grep "^TIME" gpurun_out/tp_fused2.log | cut -c1-420
timeout 200 $R --nproc-per-node 2 --master-port 29561 tools/multigpu_check.py gpurun_out/multigpu2.json > gpurun_out/multigpu2.log 2>&1; grep -E "multigpu:|FAIL" gpurun_out/multigpu2.log | head -5
timeout 120 $R --nproc-per-node 2 --master-port 29572 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench2.log 2>&1; tail -1 gpurun_out/bench2.log | cut -c1-1500
timeout 120 $R --nproc-per-node 2 --master-port 29573 bench.py --gpus 2 --steps 100 --warmup 10 --no_fused_adam > gpurun_out/bench2_unfused.log 2>&1; tail -1 gpurun_out/bench2_unfused.log | cut -c1-300
timeout 120 $R --nproc-per-node 2 --master-port 29574 bench.py --gpus 2 --steps 100 --warmup 10 --no_fused_adam --bucket_layout layers > gpurun_out/bench2_unfused_layers.log 2>&1; tail -1 gpurun_out/bench2_unfused_layers.log | cut -c1-300
timeout 200 python tensor_parallel_train.py --world_size 2 --epochs 3 --sample_size 8192 --logs_dir gpurun_out/logs_tp2 > gpurun_out/tp2.log 2>&1; grep -E "Epoch \[|Error|completed in|Traceback|probe" gpurun_out/tp2.log | head -8
timeout 200 python layer_model_parallel_train.py --world_size 2 --epochs 3 --sample_size 8192 --logs_dir gpurun_out/logs_pp2 > gpurun_out/pp2.log 2>&1; grep -E "Epoch \[|Error|completed in|Traceback|capture failed" gpurun_out/pp2.log | head -8
timeout 200 python layer_model_parallel_train.py --world_size 2 --epochs 3 --sample_size 8192 --no_pp_overlap --logs_dir gpurun_out/logs_pp2_blocking > gpurun_out/pp2_blocking.log 2>&1; grep -E "Epoch \[3|Error|completed in|Traceback" gpurun_out/pp2_blocking.log | head -4
timeout 200 python data_parallel_train.py --world_size 2 --epochs 3 --sample_size 16384 --logs_dir gpurun_out/logs_dp2 > gpurun_out/dp2.log 2>&1; grep -E "Epoch \[|Error|completed in|Traceback" gpurun_out/dp2.log | head -6
python - <<'PY'
import pandas as pd, glob, json
for d in ("logs_dp2","logs_tp2","logs_pp2","logs_pp2_blocking"):
    for f in glob.glob(f"gpurun_out/{d}/combined_results_*.csv"):
        df=pd.read_csv(f); g=df.groupby("epoch")[["loss","accuracy","epoch_time","images_per_sec","compute_time","comm_time","fwd_ms","bwd_ms","allreduce_ms","p2p_ms","exposed_comm_ms"]].max(); print(d); print(g.to_string())
    for f in glob.glob(f"gpurun_out/{d}/summary_*.json"):
        s=json.load(open(f)); print({k:v for k,v in s.items() if k in ("fused_tp","library_collectives_in_step","native_fallbacks","graph","graph_error","fused_adam","bucket_algos")})
PY
