#!/bin/bash
# N-GPU parity (all modes) + the N-GPU bench line with its strong-scaling extras.  Usage: tools_gpu_distN.sh N
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
LTX_REQUIRE_P2P=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29751 tests/dist_check.py 2>&1 | grep -E "rank 0|DIST_CHECK|rror|timeout" | tail -14
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29752 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench exit $?"; grep -v "OMP\|\*\*\*" gpurun_out/bench_${N}gpu.err | tail -5
python -c "
import json; b=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
print('steps/s', b['value'], 'ms', b['ms_per_step'], 'scaling', b['scaling'], 'e2e', b['e2e']['value'], b['clocks'])
print('parity', json.dumps(b['parity']))
print('vae', b['vae']['value'], b['vae']['ms_per_decode'])
for k,v in b['extras'].items():
    if isinstance(v, list): print(k, v); continue
    kc=v.pop('kernel_classes',None); kc2=v.pop('stage2_kernel_classes',None); print(k, v)
    for q in (kc, kc2):
        if q: print('   ', {a:(round(c['ms'],2), c['launches']) for a,c in q.items()})
"
