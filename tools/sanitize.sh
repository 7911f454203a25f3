#!/bin/bash
# compute-sanitizer evidence for the kernels: memcheck and racecheck over smoke-size builds of
# every mode.  Run on a GPU box; logs go to gpurun_out/ (copy the summaries to profiles/).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  SANITIZE_N=${SANITIZE_N:-30000} timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python tools/sanitize_run.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "compute-sanitizer --tool $tool: exit code $?" | tee -a gpurun_out/sanitizer_$tool.log
  tail -4 gpurun_out/sanitizer_$tool.log
done
