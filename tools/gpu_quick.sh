#!/bin/bash
mkdir -p gpurun_out
tools/gpu_suite.sh "$@" > gpurun_out/suite.log 2>&1; grep -E "^===|passed|failed|exit|Error|error" gpurun_out/suite.log | head -40
timeout 300 python tools/attn_bench.py 2>&1 | tee gpurun_out/attn_bench.log | tail -2
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err
python -c "
import json; b=json.load(open('gpurun_out/bench.json'))
print('steps/s', b['value'], 'ms', b['ms_per_step'], 'e2e', b['e2e']['value'], 'launches', b['gpu_launches'])
print('roofline', b['roofline']['achieved'], b['roofline']['frac'])
for k,v in b['kernel_classes'].items(): print(k, v)
print('vae', b['vae']['value'], b['vae']['ms_per_decode'], b['vae']['conv_tflops'])
"
