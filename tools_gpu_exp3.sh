#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -15
echo "=== epilogue on"; timeout 600 python tools_gemm_bench.py --bf16 2>&1 | tee gpurun_out/gemm_bench_4cta.log
echo "=== epilogue off"; LTX_GEMM_DEBUG=1 timeout 600 python tools_gemm_bench.py --bf16 2>&1 | tee gpurun_out/gemm_bench_4cta_noepi.log
