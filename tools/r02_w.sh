#!/bin/bash
# final lines of config 3 and config 4 with the RMQ chaining of the last commit (full size, reference run and GAF check of every read)
cd "$(dirname "$0")/.."
O=gpurun_out/r02w; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
timeout 700 python bench.py --workload c4 --steps 2 --warmup 1 > $O/bench_c4.json 2> $O/bench_c4.err
timeout 900 python bench.py --steps 3 --warmup 2 > $O/bench_c3.json 2> $O/bench_c3.err
tail -2 $O/pytest_gpu.txt
for f in bench_c4 bench_c3; do python - $O/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1].split("/")[-1], "value %.3f e2e %.3f cpu %.4f" % (d["value"], d["e2e"]["value"], d["cpu_baseline"]["value"]), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()}, d.get("parity_check"))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
