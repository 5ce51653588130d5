#!/bin/bash
# First hardware run of everything written after round 2's GPU budget was spent (one GPU, ~6 min):
#   1. the `late` test tier as ordinary (gating) tests, most certain first      -> gpurun_out/late_tier.log
#   2. the HZPERF report lines (persistent vs latency conv kernel vs cuDNN, MobileNetV2 / ResNet-18 step times,
#      bench.py --batch 512 with both kernel families)                            -> gpurun_out/hzperf.txt
#   3. conv roofline with the persistent kernels next to the one-tile-per-CTA kernels -> gpurun_out/conv_roofline_persist.json
#   4. the HZPERF ncu lines come from tests/test_gpu_ncu_report.py (Nsight Compute over one eager step, per-kernel shares)
# Usage: gpurun --timeout 900 -- bash tools/late_suite.sh
mkdir -p gpurun_out
HZ_LATE_STRICT=1 HZ_LATE_BUDGET_S=3600 timeout 800 python -m pytest tests -m "gpu and late and not multigpu" -q -rA -W default 2>&1 | tee gpurun_out/late_tier.log | tail -40
grep -h "HZPERF" gpurun_out/late_tier.log | sed 's/.*HZPERF/HZPERF/' | sort -u > gpurun_out/hzperf.txt
for sec in steps handoff bench conv bigbatch; do timeout 200 python tools/perf_probe.py $sec 2>/dev/null | grep HZPERF >> gpurun_out/hzperf.txt; done
cat gpurun_out/hzperf.txt
for v in 0 1; do HZ_BN_BWD_IN_DGRAD=$v timeout 120 python bench.py --gpus 1 --steps 200 --warmup 20 2>/dev/null | tail -1 | cut -c1-300 | sed "s/^/b64 HZ_BN_BWD_IN_DGRAD=$v: /"; done | tee -a gpurun_out/hzperf.txt
for v in 0 1; do HZ_CONV_PERSIST=$v timeout 120 python bench.py --gpus 1 --batch 512 --steps 50 --warmup 10 2>/dev/null | tail -1 | cut -c1-300 | sed "s/^/b512 HZ_CONV_PERSIST=$v: /"; done | tee -a gpurun_out/hzperf.txt
HZ_CONV_PERSIST=1 timeout 300 python tools/conv_roofline.py gpurun_out/conv_roofline_persist.json 512 4096 > gpurun_out/conv_roofline_persist.log 2>&1
tail -8 gpurun_out/conv_roofline_persist.log | cut -c1-400
