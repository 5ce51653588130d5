#!/bin/bash
# One-GPU evidence run: (1) every kernel launch of one eager training step with its device time, (2) a full ncu capture of
# the hot kernels (tcgen05 conv fwd / dgrad / wgrad, fused Adam, BN kernels, head), (3) compute-sanitizer passes over the
# single-kernel numerics tests.  Outputs under gpurun_out/ (reports: gpurun_out/ncu/*.ncu-rep); summaries are extracted on
# the CPU box with tools/ncu_summary.py and committed under profiles/.
mkdir -p gpurun_out/ncu
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 140 --csv --log-file gpurun_out/launches_step.csv \
    python bench.py --steps 1 --warmup 3 --no_graph > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_step.csv)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"igemm_kernel|wgrad_kernel" -s 60 -c 9 -f -o gpurun_out/ncu/conv \
    python bench.py --steps 1 --warmup 3 --no_graph > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"adam_kernel|bn_act_fwd_kernel|bn_act_bwd_apply_kernel|channel_reduce_kernel|head_sample_kernel|maxpool" -s 20 -c 8 -f -o gpurun_out/ncu/elementwise \
    python bench.py --steps 1 --warmup 3 --no_graph > gpurun_out/ncu_elem.log 2>&1; echo "ncu elementwise rc=$?"
for tool in memcheck synccheck racecheck; do
  HZ_PDL=0 timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 300 \
      python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tp.py -m gpu -x -q -k "(conv_fwd_tcgen05 or conv_dgrad or conv_wgrad or narrow or bn_act or maxpool or head or adam_kernel or stem) and not virtual and not fused" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1) | $(tail -1 gpurun_out/sanitize_$tool.log)"
done
