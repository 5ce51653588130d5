#!/bin/bash
# The 8-GPU evaluation suite (one `gpurun --gpus 8` call): collective checks, bench at 8/4/2/1 GPUs (+ NCCL variant),
# fused-TP checks, the three trainer CLIs, the hybrid mesh and the main.py sweep (BASELINE.json configs 2-5).
mkdir -p gpurun_out; R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $R --nproc-per-node 8 --master-port 29561 tools/multigpu_check.py gpurun_out/multigpu8.json > gpurun_out/multigpu8.log 2>&1; grep -E "multigpu:|FAIL" gpurun_out/multigpu8.log | head -5
for n in 8 4 2; do timeout 120 $R --nproc-per-node $n --master-port 2957$n bench.py --gpus $n --steps 200 --warmup 20 > gpurun_out/bench$n.log 2>&1; tail -1 gpurun_out/bench$n.log | cut -c1-200; done
timeout 120 python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/bench1.log 2>&1; tail -1 gpurun_out/bench1.log | cut -c1-200
timeout 120 $R --nproc-per-node 8 --master-port 29581 bench.py --gpus 8 --steps 200 --warmup 20 --allreduce nccl > gpurun_out/bench8_nccl.log 2>&1; grep '"metric"' gpurun_out/bench8_nccl.log | cut -c1-200
timeout 150 $R --nproc-per-node 8 --master-port 29562 tools/tp_fused_check.py gpurun_out/tp_fused8.json > gpurun_out/tp_fused8.log 2>&1; grep -E "FAIL|tp_fused:" gpurun_out/tp_fused8.log | head -5
timeout 150 python data_parallel_train.py --world_size 8 --epochs 3 --sample_size 50000 --logs_dir gpurun_out/logs_dp8 > gpurun_out/dp8.log 2>&1; grep -E "Epoch \[|Error|completed in" gpurun_out/dp8.log | head -5
timeout 150 python tensor_parallel_train.py --world_size 8 --epochs 2 --sample_size 8192 --logs_dir gpurun_out/logs_tp8 > gpurun_out/tp8.log 2>&1; grep -E "Epoch \[2|Error|completed in|Traceback" gpurun_out/tp8.log | head -3
timeout 150 python layer_model_parallel_train.py --world_size 4 --epochs 3 --sample_size 8192 --logs_dir gpurun_out/logs_pp4 > gpurun_out/pp4.log 2>&1; grep -E "Epoch \[|Error|completed in|Traceback|capture failed" gpurun_out/pp4.log | head -5
timeout 200 python hybrid_parallel_train.py --world_size 8 --dp_replicas 2 --inner layer --epochs 2 --sample_size 16384 --logs_dir gpurun_out/logs_hybrid8 > gpurun_out/hybrid8.log 2>&1; grep -E "Epoch \[2|Error|completed in|Traceback" gpurun_out/hybrid8.log | head -3
timeout 600 python main.py --sample_sizes 1000 10000 50000 --world_size 8 --epochs 2 --output_dir gpurun_out/benchmark_results_8gpu > gpurun_out/main8.log 2>&1; tail -2 gpurun_out/main8.log
python - <<'PY'
import pandas as pd, glob
for d in ("logs_dp8","logs_tp8","logs_pp4"):
    for f in glob.glob(f"gpurun_out/{d}/combined_results_*.csv"):
        df=pd.read_csv(f); g=df.groupby("epoch")[["loss","accuracy","epoch_time","images_per_sec"]].max(); print(d); print(g.to_string())
PY
