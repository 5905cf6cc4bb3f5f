#!/bin/bash
# round 2, session 2, second GPU call: tests on the candidate build, then A/B/C of k_wfa_big (two cells in flight at 20 / 16 / 24 warps per SM)
# and k_gwfa (shared-memory arena / none / none at 24 warps per SM)
cd "$(dirname "$0")/.."
O=gpurun_out/r02n; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
B="--reads 40000 --steps 2 --warmup 2 --no-cpu"
timeout 400 python bench.py $B > $O/c3_A.json 2> $O/c3_A.err
MGB_LIB=tools/ab/libmgb200_B.so timeout 400 python bench.py $B > $O/c3_B.json 2> $O/c3_B.err
MGB_LIB=tools/ab/libmgb200_C.so timeout 400 python bench.py $B > $O/c3_C.json 2> $O/c3_C.err
MGB_LIB=tools/ab/libmgb200_C.so timeout 300 python -m pytest tests -m gpu -q -x -k "tiers or c2_mt or c3_sv or c4_asm or fallback or label" > $O/pytest_C.log 2>&1; echo "pytest rc=$?" >> $O/pytest_C.log
tail -2 $O/pytest.log; tail -2 $O/pytest_C.log
for f in c3_A c3_B c3_C; do python - $O/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1].split("/")[-1], "value %.3f e2e %.3f" % (d["value"], d["e2e"]["value"]), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()}, d.get("parity_check"))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
