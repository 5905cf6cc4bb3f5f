#!/bin/bash
# round 2, session 2, fifth GPU call: resident warps of the arena-heavy kernels (their working sets against L2) and the mini-batch size
cd "$(dirname "$0")/.."
O=gpurun_out/r02q; mkdir -p $O
B="--reads 40000 --steps 2 --warmup 2 --no-cpu"
timeout 300 python bench.py $B > $O/c3_p0.json 2> $O/c3_p0.err
MGB_PARAMS=mb0=4 timeout 300 python bench.py $B > $O/c3_p1.json 2> $O/c3_p1.err
MGB_PARAMS=mb0=6 timeout 300 python bench.py $B > $O/c3_p2.json 2> $O/c3_p2.err
MGB_PARAMS=mb5=4,mb9=2,mb2=4 timeout 300 python bench.py $B > $O/c3_p3.json 2> $O/c3_p3.err
MGB_PARAMS=mb5=6,mb9=3 timeout 300 python bench.py $B > $O/c3_p4.json 2> $O/c3_p4.err
timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu > $O/c3_full_mb400.json 2> $O/c3_full_mb400.err
timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu --mini-batch 800000000 > $O/c3_full_mb800.json 2> $O/c3_full_mb800.err
for f in c3_p0 c3_p1 c3_p2 c3_p3 c3_p4 c3_full_mb400 c3_full_mb800; do python - $O/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1].split("/")[-1], "value %.3f e2e %.3f" % (d["value"], d["e2e"]["value"]), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()})
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
