#!/bin/bash
# config 4 (asm preset): where k_chain spends its time
cd "$(dirname "$0")/.."
O=gpurun_out/r02u; mkdir -p $O
BENCH_ARGS="--workload c4 --reads 8000" timeout 500 tools/profile_kernels.sh r02c4 k_chain k_seed > $O/prof.log 2>&1
tail -3 $O/prof.log
