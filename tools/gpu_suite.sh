#!/bin/bash
# Runs each GPU test file in its own process (a kernel trap poisons the CUDA context of that process only).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in ${@:-gemm attention elementwise conv dit vae quant}; do
  echo "=== $f"
  timeout 600 python -m pytest tests/test_gpu_$f.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit $?"
  tail -5 gpurun_out/test_$f.log
done
