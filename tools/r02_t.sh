#!/bin/bash
# round 2, session 2, eighth GPU call: tier 3 with three and four cells per lane in flight
cd "$(dirname "$0")/.."
O=gpurun_out/r02t; mkdir -p $O
B="--reads 40000 --steps 2 --warmup 2 --no-cpu"
for v in c3 c4; do MGB_LIB=tools/ab/libmgb200_$v.so timeout 300 python bench.py $B > $O/c3_$v.json 2> $O/c3_$v.err; done
MGB_LIB=tools/ab/libmgb200_c4.so timeout 300 python -m pytest tests -m gpu -q -x -k "tiers or fallback or c3_sv or larger" > $O/pytest_c4.log 2>&1; echo "pytest rc=$?" >> $O/pytest_c4.log
tail -2 $O/pytest_c4.log
for f in c3_c3 c3_c4; do python - $O/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1].split("/")[-1], "value %.3f e2e %.3f" % (d["value"], d["e2e"]["value"]), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()})
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
