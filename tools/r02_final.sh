#!/bin/bash
# Round 2 evidence set (run on the GPU box through gpurun): tests, the bench lines of every workload, the reference arm,
# the ncu launch list and one `ncu --set full` digest per kernel.  Everything lands in gpurun_out/r02/ and gpurun_out/prof_r02_*.
cd "$(dirname "$0")/.."
O=gpurun_out/r02; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
timeout 900 python bench.py --steps 3 --warmup 2 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_c3_reference.json 2> $O/bench_c3_reference.err
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 > $O/bench_c4.json 2> $O/bench_c4.err
timeout 900 python bench.py --workload c5 --steps 2 --warmup 1 --check 10000 > $O/bench_c5.json 2> $O/bench_c5.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3_20k.csv python bench.py --reads 20000 --mini-batch 400000000 --steps 1 --warmup 1 --no-cpu --pipe 1 > $O/launches.log 2>&1
BENCH_ARGS="--reads 20000" tools/profile_kernels.sh r02 k_gwfa k_wfa_big k_wfa_mid k_chain k_chain_rescue k_seed k_wfa_small k_finish k_gchain k_gchain_gen k_gc_labels > $O/prof.log 2>&1
BENCH_ARGS="--workload c2" tools/profile_kernels.sh r02c2 k_chain k_chain_rescue >> $O/prof.log 2>&1
tail -3 $O/pytest_gpu.txt
