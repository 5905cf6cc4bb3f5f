#!/bin/bash
# round 2, session 2, fourth GPU call: k_gwfa with its three phases out of line at 20 / 24 / 32 warps per SM (96 / 80 / 64 registers)
cd "$(dirname "$0")/.."
O=gpurun_out/r02p; mkdir -p $O
B="--reads 40000 --steps 2 --warmup 2 --no-cpu"
timeout 400 python bench.py $B > $O/c3_main.json 2> $O/c3_main.err
for v in g5 g6 g8; do MGB_LIB=tools/ab/libmgb200_$v.so timeout 400 python bench.py $B > $O/c3_$v.json 2> $O/c3_$v.err; done
MGB_LIB=tools/ab/libmgb200_g8.so timeout 400 python -m pytest tests -m gpu -q -x -k "c2_mt or c3_sv or c4_asm or larger or struct or label or concurrent" > $O/pytest_g8.log 2>&1; echo "pytest rc=$?" >> $O/pytest_g8.log
tail -2 $O/pytest_g8.log
for f in c3_main c3_g5 c3_g6 c3_g8; do python - $O/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1].split("/")[-1], "value %.3f e2e %.3f" % (d["value"], d["e2e"]["value"]), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()})
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
