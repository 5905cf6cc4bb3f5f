timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { MGB_PARAMS=$1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --check 3000 > gpurun_out/x.json 2> gpurun_out/x.err; tail -2 gpurun_out/x.err; python -c "
import json;d=json.load(open('gpurun_out/x.json'));print('$1', round(d['value'],4), round(d['ms_per_step'],1), round(d['e2e']['value'],4), round(d['e2e']['ms_per_step'],1), d['kernel_ms'], d['stage_ms_per_step']['wfa_jobs_tier2'], d['stage_ms_per_step']['wfa_jobs_tier3'], d['parity_check']['identical'])"; }
run tier_learn=1
run tier_learn=0
timeout 600 python bench.py --workload c3 --reads 20000 --steps 2 --warmup 1 --no-cpu --check 2000 > gpurun_out/y.json 2> gpurun_out/y.err; tail -2 gpurun_out/y.err; python -c "
import json;d=json.load(open('gpurun_out/y.json'));print('c3', round(d['value'],4), round(d['ms_per_step'],1), round(d['e2e']['value'],4), round(d['e2e']['ms_per_step'],1), d['kernel_ms'], d['parity_check'])"
