#!/bin/bash
# Round 2, final evidence set of the second session (run on the GPU box through gpurun): tests, the bench lines of the workloads, the ncu
# launch list and one `ncu --set full` digest per kernel of the final code.  Everything lands in gpurun_out/r02f/ and gpurun_out/prof_r02*.
cd "$(dirname "$0")/.."
O=gpurun_out/r02f; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
timeout 900 python bench.py --steps 3 --warmup 2 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 400 python bench.py --workload c2 --steps 10 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3_20k.csv python bench.py --reads 20000 --mini-batch 400000000 --steps 1 --warmup 1 --no-cpu --pipe 1 > $O/launches.log 2>&1
BENCH_ARGS="--reads 20000" timeout 900 tools/profile_kernels.sh r02 k_wfa_mid k_wfa_big k_gwfa k_wfa_small k_seed k_finish k_gchain k_gchain_gen k_gc_labels > $O/prof.log 2>&1
BENCH_ARGS="--reads 20000" timeout 300 tools/profile_kernels.sh r02c3 k_chain k_chain_rescue >> $O/prof.log 2>&1
BENCH_ARGS="--workload c2" timeout 300 tools/profile_kernels.sh r02c2 k_chain k_chain_rescue >> $O/prof.log 2>&1
timeout 700 python bench.py --workload c5 --steps 2 --warmup 1 --check 10000 > $O/bench_c5.json 2> $O/bench_c5.err
timeout 700 python bench.py --workload c4 --steps 2 --warmup 1 > $O/bench_c4.json 2> $O/bench_c4.err
tail -2 $O/pytest_gpu.txt
for f in bench_c3 bench_c2 bench_c5 bench_c4; do python - $O/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1].split("/")[-1], "value %.3f e2e %.3f cpu %.4f" % (d["value"], d["e2e"]["value"], d["cpu_baseline"]["value"]), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()}, d.get("parity_check"))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
