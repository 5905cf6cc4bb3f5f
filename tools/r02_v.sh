#!/bin/bash
# RMQ chaining: two summaries per block (split at the middle of the block's query span), border scan four half-blocks at a time
cd "$(dirname "$0")/.."
O=gpurun_out/${OUT:-r02v}; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 500 python bench.py --workload c4 --reads 20000 --steps 2 --warmup 1 --no-cpu > $O/c4.json 2> $O/c4.err
timeout 300 python bench.py --reads 40000 --steps 2 --warmup 2 --no-cpu > $O/c3.json 2> $O/c3.err
tail -2 $O/pytest.log
for f in c4 c3; do python - $O/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1].split("/")[-1], "value %.3f e2e %.3f" % (d["value"], d["e2e"]["value"]), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()})
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
