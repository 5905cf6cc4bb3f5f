#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02j; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
B="--reads 40000 --steps 2 --warmup 2 --no-cpu"
timeout 600 python bench.py $B > $O/base.json 2> $O/base.err
MGB_PARAMS=mb7=2 timeout 600 python bench.py $B > $O/mb7_2.json 2> $O/mb7_2.err
MGB_PARAMS=mb7=3 timeout 600 python bench.py $B > $O/mb7_3.json 2> $O/mb7_3.err
MGB_PARAMS=mb7=1 timeout 600 python bench.py $B > $O/mb7_1.json 2> $O/mb7_1.err
MGB_PARAMS=mb8=2 timeout 600 python bench.py $B > $O/mb8_2.json 2> $O/mb8_2.err
MGB_PARAMS=mb8=6 timeout 600 python bench.py $B > $O/mb8_6.json 2> $O/mb8_6.err
BENCH_ARGS="--reads 25000" tools/profile_kernels.sh r02c3 k_chain k_chain_rescue > $O/prof.log 2>&1
BENCH_ARGS="--workload c2" tools/profile_kernels.sh r02c2 k_chain k_chain_rescue >> $O/prof.log 2>&1
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu > $O/c2.json 2> $O/c2.err
tail -3 $O/pytest.log
