#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02b; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python bench.py --no-cpu --check 1000 > $O/c2.json 2> $O/c2.err
timeout 900 python bench.py --workload c3 --reads 20000 --no-cpu --check 2000 > $O/c3.json 2> $O/c3.err
MGB_PARAMS=lab_cache=0 timeout 900 python bench.py --workload c3 --reads 20000 --no-cpu --check 0 > $O/c3_nocache.json 2> $O/c3_nocache.err
MGB_PARAMS=slots=2 timeout 900 python bench.py --workload c3 --reads 20000 --no-cpu --check 0 > $O/c3_slots2.json 2> $O/c3_slots2.err
tail -5 $O/pytest.log
