#!/bin/bash
# round 2, session 2, third GPU call: the packed-pair wavefront walk (two adjacent diagonals per lane, VIMNMX.S16x2) against the scalar one
cd "$(dirname "$0")/.."
O=gpurun_out/r02o; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
B="--reads 40000 --steps 2 --warmup 2 --no-cpu"
timeout 400 python bench.py $B > $O/c3_main.json 2> $O/c3_main.err
MGB_LIB=tools/ab/libmgb200_P4.so timeout 400 python bench.py $B > $O/c3_P4.json 2> $O/c3_P4.err
MGB_LIB=tools/ab/libmgb200_P5.so timeout 400 python bench.py $B > $O/c3_P5.json 2> $O/c3_P5.err
MGB_LIB=tools/ab/libmgb200_P5.so timeout 400 python -m pytest tests -m gpu -q -x -k "tiers or c2_mt or c3_sv or c4_asm or fallback or larger or struct or routing or short" > $O/pytest_P5.log 2>&1; echo "pytest rc=$?" >> $O/pytest_P5.log
tail -2 $O/pytest.log; tail -2 $O/pytest_P5.log
for f in c3_main c3_P4 c3_P5; do python - $O/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1].split("/")[-1], "value %.3f e2e %.3f" % (d["value"], d["e2e"]["value"]), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()}, d.get("wfa_tier_routing"))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
