#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02d; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu > $O/c3.json 2> $O/c3.err
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu --pipe 1 > $O/c3_pipe1.json 2> $O/c3_pipe1.err
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu --pipe 2 > $O/c3_pipe2.json 2> $O/c3_pipe2.err
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu > $O/c2.json 2> $O/c2.err
tail -5 $O/pytest.log
