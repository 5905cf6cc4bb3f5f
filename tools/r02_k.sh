#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02k; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python bench.py --reads 40000 --steps 2 --warmup 2 --no-cpu > $O/c3_40k.json 2> $O/c3_40k.err
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu > $O/c2.json 2> $O/c2.err
timeout 900 python bench.py --workload c4 --reads 20000 --steps 2 --warmup 1 --no-cpu > $O/c4.json 2> $O/c4.err
tail -3 $O/pytest.log
