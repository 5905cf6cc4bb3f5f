#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02c; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 900 python bench.py --steps 3 --warmup 2 > $O/c3.json 2> $O/c3.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/c3_ref.json 2> $O/c3_ref.err
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 > $O/c2.json 2> $O/c2.err
timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 > $O/c4.json 2> $O/c4.err
tail -5 $O/pytest.log
