#!/bin/bash
mkdir -p gpurun_out
./tools_gpu_suite.sh > gpurun_out/suite.log 2>&1; grep -E "^===|passed|failed|exit" gpurun_out/suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 16 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err
python -c "
import json; b=json.load(open('gpurun_out/bench.json'))
print('steps/s', b['value'], 'ms', b['ms_per_step'], 'e2e', b['e2e']['value'], 'launches', b['gpu_launches'], 'cpu', b['cpu_baseline'])
print('roofline', b['roofline']['achieved'], b['roofline']['frac'])
for k,v in b['kernel_classes'].items(): print(k, v)
print('vae', b['vae']['value'], b['vae']['ms_per_decode'], b['vae']['conv_tflops'], b['vae']['e2e_value'])
print(b['extras']); print(b['clocks'])
"
timeout 300 python tools_ncu_target.py > gpurun_out/plain_ncu_target.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 1 -c 2 -o gpurun_out/r01_gemm_ffn_in python tools_ncu_target.py > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 1 -c 1 -o gpurun_out/r01_attention python tools_ncu_target.py > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_tcgen05 -s 1 -c 1 -o gpurun_out/r01_conv3d python tools_ncu_target.py > gpurun_out/ncu_conv.log 2>&1
echo "ncu conv exit $?"
ls -la gpurun_out/*.ncu-rep
