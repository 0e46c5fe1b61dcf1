#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_quant.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -8
timeout 600 python tools_gemm_bench.py --quant 2>&1 | tee gpurun_out/gemm_q_bench.log
