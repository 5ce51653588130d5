#!/bin/bash
# compute-sanitizer passes over the GPU kernel tests: memcheck (OOB / misaligned, incl. TMA-fed
# shared memory), synccheck (barrier misuse), racecheck (shared-memory hazards of the elementwise / reduction kernels
# and of the virtual-rank peer all-reduce, whose flag protocol is exercised back to back by the test).
# One GPU; slow (10-50x) — run through `gpurun --timeout 1500 -- tools/sanitize.sh [memcheck|synccheck|racecheck]`.
# Results: gpurun_out/sanitize_<tool>.log (+ summary line per tool).
set -u
mkdir -p gpurun_out
TOOLS=${1:-"memcheck synccheck racecheck"}
# the graph-capture / whole-model tests are excluded: the sanitizer serialises kernels and the step takes minutes
SEL='conv_fwd or conv_dgrad or conv_wgrad or bn_act or maxpool or head or adam or peer_allreduce or stem'
for tool in $TOOLS; do
  log=gpurun_out/sanitize_$tool.log
  HZ_PDL=0 timeout 1400 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 300 \
      python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "$SEL" > $log 2>&1
  rc=$?
  echo "$tool rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) $(tail -1 $log)"
done
