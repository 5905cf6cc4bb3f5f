#!/bin/bash
# last GPU call of the round: the digit walk of the exact radix sort (k_seed, k_chain, k_chain_rescue)
cd "$(dirname "$0")/.."
O=gpurun_out/r02y; mkdir -p $O
timeout 100 python bench.py --reads 40000 --steps 2 --warmup 2 --no-cpu > $O/c3.json 2> $O/c3.err
timeout 120 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -2 $O/pytest.log
python - $O/c3.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print("value %.3f e2e %.3f" % (d["value"], d["e2e"]["value"]), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()})
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
