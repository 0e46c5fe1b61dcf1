#!/bin/bash
# round-2 first GPU pass: whole GPU suite (per file, own process), small-M GEMM microbench, the self-validating bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in gemm quant attention elementwise conv dit vae golden av encoder_upscaler baseline_shapes; do
  echo "=== $f"
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit $?"
  tail -4 gpurun_out/test_$f.log
done
timeout 300 python tools/gemm_bench.py --small > gpurun_out/gemm_small.log 2>&1; echo "gemm_small exit $?"; cat gpurun_out/gemm_small.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench exit $?"; tail -3 gpurun_out/bench_r02a.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_r02a.json'))
print('steps/s', b['value'], 'ms', b['ms_per_step'], 'e2e', b['e2e']['value'], 'launches', b['gpu_launches'], 'gap', b['step_minus_class_sum_ms'])
print('parity', json.dumps(b['parity']))
print('roofline', b['roofline']['achieved'], b['roofline']['frac'])
for k,v in b['kernel_classes'].items(): print(k, v)
print('vae', b['vae']['value'], b['vae']['ms_per_decode'], b['vae']['conv_tflops'])
print('cpu', b['cpu_baseline'])
for k,v in b['extras'].items(): print(k, {kk:vv for kk,vv in v.items() if 'kernel_classes' not in kk})
PY
