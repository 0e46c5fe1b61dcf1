#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_quant.py tests/test_gpu_dit.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 900 python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err
python -c "
import json; b=json.load(open('gpurun_out/bench.json'))
print('steps/s', b['value'], 'ms', b['ms_per_step'], 'e2e', b['e2e']['value'], 'launches', b['gpu_launches'], b['clocks'])
print('roofline', b['roofline']['achieved'], b['roofline']['frac'])
for k,v in b['kernel_classes'].items(): print(k, v)
print('vae', b['vae']['value'], b['vae']['ms_per_decode'], b['vae']['conv_tflops'])
print(b['extras'])
"
