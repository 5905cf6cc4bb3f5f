#!/bin/bash
# round 2, session 2, sixth and seventh GPU call: counters summed per block in shared memory, per-warp slices of the CIGAR pool, WFA jobs taken four at a time
cd "$(dirname "$0")/.."
O=gpurun_out/${OUT:-r02r}; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
B="--reads 40000 --steps 2 --warmup 2 --no-cpu"
timeout 300 python bench.py $B > $O/c3_new.json 2> $O/c3_new.err
timeout 300 python bench.py $B > $O/c3_new2.json 2> $O/c3_new2.err
tail -2 $O/pytest.log
for f in c3_new c3_new2; do python - $O/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1].split("/")[-1], "value %.3f e2e %.3f" % (d["value"], d["e2e"]["value"]), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()})
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
