#!/bin/bash
# Run on the GPU box (through gpurun): the evidence set of one round.  Usage: tools/profile_round.sh ROUND
#   gpurun_out/<round>_pytest_gpu.txt      parity tests through the C ABI
#   gpurun_out/<round>_bench_1gpu.json     the bench line (CUDA-event timing, never under a profiler)
#   gpurun_out/<round>_launches.csv        ncu launch list (gpu__time_duration.sum, --clock-control none) of the same command
#   gpurun_out/prof_<round>_<kernel>.*     `ncu --set full` digests of the kernels named after ROUND (default: k_chain k_wfa_mid)
round=$1; shift
kernels=${@:-k_chain k_wfa_mid}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${round}_pytest_gpu.txt
cat gpurun_out/${round}_pytest_gpu.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${round}_bench_1gpu.json 2> gpurun_out/${round}_bench_1gpu.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${round}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${round}_launches.log 2>&1
tools/profile_kernels.sh $round $kernels
