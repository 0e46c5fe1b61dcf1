#!/bin/bash
mkdir -p gpurun_out
for f in dit golden quant; do
  echo "=== $f"
  timeout 600 python -m pytest tests/test_gpu_$f.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit $?"; tail -6 gpurun_out/test_$f.log
done
for b in 1 0; do
  LTX_BATCHED_CFG=$b timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity --no-cfg5 > gpurun_out/bench_bcfg$b.json 2> gpurun_out/bench_bcfg$b.err; echo "bench LTX_BATCHED_CFG=$b exit $?"; tail -2 gpurun_out/bench_bcfg$b.err
  python - <<PY
import json
b=json.load(open('gpurun_out/bench_bcfg$b.json'))
print('LTX_BATCHED_CFG=$b steps/s', b['value'], 'ms', b['ms_per_step'], 'guided', b['extras']['guided_cfg3']['ms_per_step'])
PY
done
