#!/bin/bash
# Run on the GPU box (through gpurun): one `ncu --set full` capture per named kernel of the bench workload, plus the
# per-source-line digest of each.  Usage: tools/profile_kernels.sh ROUND k_wfa_mid k_gwfa ...
# Outputs: gpurun_out/prof_<round>_<kernel>.ncu-rep, .lines.txt (hot source lines), .metrics.csv (raw page)
set -u
round=$1; shift
mkdir -p gpurun_out
lib=minigraph_b200/libmgb200.so
cuobjdump -xelf all $lib > /dev/null 2>&1
cubin=$(ls -t *.cubin | head -1)
nvdisasm -g -c $cubin > gpurun_out/dis.txt 2>/dev/null
for k in "$@"; do
	out=gpurun_out/prof_${round}_$k
	timeout 900 ncu --set full --clock-control none --import-source on -k $k -c 1 -f -o $out python bench.py --steps 1 --warmup 1 --no-cpu --pipe 1 ${BENCH_ARGS:-} > $out.log 2>&1
	ncu -i $out.ncu-rep --page source --csv > $out.sass.csv 2>/dev/null
	ncu -i $out.ncu-rep --page raw --csv > $out.metrics.csv 2>/dev/null
	mangled=$(grep -o "_Z[0-9]*${k}10LaunchArgs" gpurun_out/dis.txt | head -1)
	python tools/ncu_lines.py $out.sass.csv gpurun_out/dis.txt $mangled 45 > $out.lines.txt 2>&1
	rm -f $out.sass.csv
	[ -n "${KEEP_REP:-}" ] || rm -f $out.ncu-rep   # gpurun brings back at most 64 MiB
done
rm -f *.cubin gpurun_out/dis.txt
