#!/bin/bash
# ncu evidence for every hot kernel class: plain run first, then one `--set full` capture per kernel, then the launch list.
mkdir -p gpurun_out
R=${1:-r02c}
timeout 300 python tools/ncu_target.py > gpurun_out/plain_ncu_target.log 2>&1 || { echo "plain target failed"; tail -5 gpurun_out/plain_ncu_target.log; exit 1; }
cap() {  # name regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/${R}_$1 python tools/ncu_target.py > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 exit $?"
}
cap gemm_pair_ffn_in gemm_bf16_2cta 2
cap gemm_swapab_splitk gemm_swapab_2cta 1
cap attention attention_fwd 1
cap conv3d_pair conv3d_pair_tcgen05 1
cap rmsnorm_mod rmsnorm_mod_stream 1
cap qknorm_rope qknorm_rope_stream 1
cap vae_prep vae_prep_kernel 1
ls -la gpurun_out/*.ncu-rep
export LTX_GRAPH=0   # the launch list is of the eager step (a replayed graph issues the same kernels)
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-cfg5 > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-cfg5 > gpurun_out/ncu.log 2>&1
echo "ncu launches exit $?"; tail -2 gpurun_out/ncu.log
