#!/bin/bash
# round 2, session 2, first GPU call: tests, then A/B of the builds under tools/ab/ (bench lines with per-kernel times), ncu of the two kernels that changed
cd "$(dirname "$0")/.."
O=gpurun_out/r02m; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
B="--reads 40000 --steps 2 --warmup 2 --no-cpu"
timeout 400 python bench.py $B > $O/c3_new.json 2> $O/c3_new.err
MGB_PARAMS=mb7=4 timeout 400 python bench.py $B > $O/c3_new_mb4.json 2> $O/c3_new_mb4.err
MGB_LIB=tools/ab/libmgb200_c2.so timeout 400 python bench.py $B > $O/c3_c2.json 2> $O/c3_c2.err
MGB_LIB=tools/ab/libmgb200_base.so timeout 400 python bench.py $B > $O/c3_base.json 2> $O/c3_base.err
MGB_LIB=tools/ab/libmgb200_c2.so timeout 300 python -m pytest tests -m gpu -q -x -k "tiers or c2_mt or c3_sv or c4_asm or fallback" > $O/pytest_c2.log 2>&1; echo "pytest rc=$?" >> $O/pytest_c2.log
BENCH_ARGS="--reads 20000" timeout 500 tools/profile_kernels.sh r02m k_gwfa k_wfa_big > $O/prof.log 2>&1
timeout 700 python bench.py --steps 3 --warmup 2 > $O/bench_c3_full.json 2> $O/bench_c3_full.err
tail -2 $O/pytest.log; tail -2 $O/pytest_c2.log
for f in c3_new c3_new_mb4 c3_c2 c3_base bench_c3_full; do python - $O/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1].split("/")[-1], "value %.3f e2e %.3f" % (d["value"], d["e2e"]["value"]), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()}, d.get("parity_check"))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
