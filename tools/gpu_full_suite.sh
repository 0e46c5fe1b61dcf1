#!/bin/bash
# the driver's round-end sequence on one GPU: the whole GPU suite in one pytest process, then smoke()
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/full_suite.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/full_suite.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
