#!/bin/bash
# round-2 third GPU pass: super-tile GEMM tests + sweep, Ulysses q|k|v merge is exercised by the 2-GPU pass later
mkdir -p gpurun_out
for f in gemm dit; do
  echo "=== $f"
  timeout 600 python -m pytest tests/test_gpu_$f.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit $?"; tail -4 gpurun_out/test_$f.log
done
timeout 400 python tools/gemm_bench.py --bf16 2>&1 | tee gpurun_out/gemm_super_sweep.log
for sm in 1 0; do
  LTX_GEMM_SUPER=$sm timeout 600 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-parity --no-cfg5 > gpurun_out/bench_super$sm.json 2> gpurun_out/bench_super$sm.err; echo "bench LTX_GEMM_SUPER=$sm exit $?"; tail -2 gpurun_out/bench_super$sm.err
  python - <<PY
import json
b=json.load(open('gpurun_out/bench_super$sm.json'))
print('LTX_GEMM_SUPER=$sm steps/s', b['value'], 'ms', b['ms_per_step'], 'e2e', b['e2e']['value'], 'gap', b['step_minus_class_sum_ms'], 'roofline', b['roofline']['achieved'], b['roofline']['frac'])
for k,v in b['kernel_classes'].items(): print('   ', k, v['ms'], v['launches'])
print('   guided', b['extras']['guided_cfg3']['ms_per_step'], 'qint8', b['extras']['qint8'].get('ms_per_step'), 'av', b['extras']['av_dual_forward'].get('ms_per_forward'))
PY
done
