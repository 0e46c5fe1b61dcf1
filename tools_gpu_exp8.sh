#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_vae.py tests/test_loader.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err
python -c "
import json; b=json.load(open('gpurun_out/bench.json'))
print('steps/s', b['value'], 'ms', b['ms_per_step'])
v=b['vae']; print('vae', v['value'], v['ms_per_decode'], v['conv_tflops'], v['kernel_classes'])
print(b['extras']['vae_121f'])
"
