#!/bin/bash
# round-2 second GPU pass: swap-AB split-K rewrite + captured-step tests, small-M microbench, quick bench
mkdir -p gpurun_out
for f in gemm dit; do
  echo "=== $f"
  timeout 600 python -m pytest tests/test_gpu_$f.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit $?"; tail -4 gpurun_out/test_$f.log
done
timeout 280 python tools/gemm_bench.py --small 2>&1 | grep -v "M= 768\|M=  96" | tee gpurun_out/gemm_small3.log
for g in 1 0; do
  LTX_GRAPH=$g timeout 600 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-parity --no-cfg5 > gpurun_out/bench_graph$g.json 2> gpurun_out/bench_graph$g.err; echo "bench LTX_GRAPH=$g exit $?"; tail -2 gpurun_out/bench_graph$g.err
  python - <<PY
import json
b=json.load(open('gpurun_out/bench_graph$g.json'))
print('LTX_GRAPH=$g steps/s', b['value'], 'ms', b['ms_per_step'], 'e2e', b['e2e']['value'], b['e2e']['ms_per_step'], 'launches', b['gpu_launches'], 'gap', b['step_minus_class_sum_ms'])
print('   guided', b['extras']['guided_cfg3']['ms_per_step'], 'qint8', b['extras']['qint8'].get('ms_per_step'), 'vae', b['vae']['ms_per_decode'], 'vae121', b['extras']['vae_121f']['ms_per_decode'])
PY
done
