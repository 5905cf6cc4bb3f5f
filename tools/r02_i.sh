#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02i; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
B="--reads 40000 --steps 3 --warmup 2 --no-cpu"
timeout 600 python bench.py $B --pipe 3 > $O/p3.json 2> $O/p3.err
timeout 600 python bench.py $B --pipe 2 > $O/p2.json 2> $O/p2.err
timeout 600 python bench.py $B --pipe 3 --mini-batch 800000000 > $O/p3_mb800.json 2> $O/p3_mb800.err
MGB_PARAMS=gpu_lock=0 timeout 600 python bench.py $B --pipe 3 > $O/p3_nolock.json 2> $O/p3_nolock.err
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu > $O/c2.json 2> $O/c2.err
tail -5 $O/pytest.log
