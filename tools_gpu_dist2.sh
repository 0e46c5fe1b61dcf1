#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m pytest tests/test_gpu_attention.py tests/test_gpu_elementwise.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -2
for p2p in 1 0; do
echo "=== LTX_P2P=$p2p"
LTX_P2P=$p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2973$p2p tests/dist_check.py sp 2>&1 | grep -E "rank|DIST_CHECK|rror|timeout" | tail -12
done
for p2p in 1 0; do
LTX_P2P=$p2p timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2974$p2p bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu_p2p$p2p.json 2> gpurun_out/bench_${N}gpu.err; echo "bench exit $?"; tail -3 gpurun_out/bench_${N}gpu.err | grep -v OMP
python -c "
import json; b=json.loads(open('gpurun_out/bench_${N}gpu_p2p$p2p.json').read().strip().splitlines()[-1])
print('P2P=$p2p steps/s', b['value'], 'ms', b['ms_per_step'], 'e2e', b['e2e']['value'])
u=b['extras']['ulysses']; print('ulysses ms', u['ms_per_step'], {k:(round(v['ms'],2), v['launches']) for k,v in u['kernel_classes'].items()})
print({k:v for k,v in b['extras'].items() if k!='ulysses'})
"
done
