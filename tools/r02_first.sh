#!/bin/bash
# round 2, first GPU call: the whole GPU suite with every opt-in kernel, then each switch timed alone
cd "$(dirname "$0")/.."
O=gpurun_out/r02a; mkdir -p $O
MGB_TEST_GEN_V2=1 MGB_TEST_CHAIN_V2=1 MGB_TEST_FIN_V2=1 MGB_TEST_SEED_V2=1 MGB_TEST_WFA_V2=1 MGB_TEST_CTA=1 \
  timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for p in "" wfa_v2=1 seed_v2=1 fin_v2=1 chain_v2=1 gen_v2=1 cta_len=1500 wfa_v2=1,seed_v2=1,fin_v2=1,chain_v2=1,gen_v2=1; do
  n=$(echo "$p" | tr ',=' '__'); [ -z "$n" ] && n=default
  MGB_PARAMS=$p timeout 600 python bench.py --no-cpu --check 1000 > $O/c2_$n.json 2> $O/c2_$n.err
done
for p in "" wfa_v2=1,seed_v2=1,fin_v2=1,chain_v2=1,gen_v2=1; do
  n=$(echo "$p" | tr ',=' '__'); [ -z "$n" ] && n=default
  MGB_PARAMS=$p timeout 900 python bench.py --workload c3 --reads 20000 --no-cpu --check 2000 > $O/c3_$n.json 2> $O/c3_$n.err
done
nproc > $O/nproc.txt; lscpu | head -20 >> $O/nproc.txt
tail -3 $O/pytest.log
