#!/bin/bash
# run with gpurun --gpus N: the torchrun path of bench.py at N ranks and the MGB_DEVICES tests on real second devices
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out/r02_n$N; mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 ${BENCH_EXTRA:-} > $O/bench_c3.json 2> $O/bench_c3.err
timeout 600 python -m pytest tests -m gpu -q -k "several or dropin" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log; tail -c 600 $O/bench_c3.json
