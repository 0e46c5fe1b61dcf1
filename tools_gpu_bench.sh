#!/bin/bash
# tests + smoke + bench + (optional) ncu launch list.  Usage: tools_gpu_bench.sh [steps] [ncu]
mkdir -p gpurun_out
./tools_gpu_suite.sh > gpurun_out/suite.log 2>&1; grep -E "^===|passed|failed|exit" gpurun_out/suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
STEPS=${1:-16}
timeout 900 python bench.py --steps $STEPS --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$2" = "ncu" ]; then
  timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
  echo "ncu exit $?"; tail -2 gpurun_out/ncu.log
fi
