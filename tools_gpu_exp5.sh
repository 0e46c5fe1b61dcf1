#!/bin/bash
mkdir -p gpurun_out
./tools_gpu_suite.sh > gpurun_out/suite.log 2>&1; grep -E "^===|passed|failed|exit|Error|error" gpurun_out/suite.log | head -40
for r in 4 2; do
LTX_ROWS_PER_CTA=$r timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rows$r.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err
python -c "
import json; b=json.load(open('gpurun_out/bench_rows$r.json'))
print('ROWS=$r steps/s', b['value'], 'ms', b['ms_per_step'], 'e2e', b['e2e']['value'], 'launches', b['gpu_launches'], b['clocks'])
for k,v in b['kernel_classes'].items(): print(k, v)
print('vae', b['vae']['value'], b['vae']['ms_per_decode'], b['vae']['conv_tflops'])
print(b['extras'])
"
done
