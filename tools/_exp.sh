run() { MGB_PARAMS=$1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --check 2000 > gpurun_out/x.json 2> gpurun_out/x.err; tail -2 gpurun_out/x.err; python -c "
import json;d=json.load(open('gpurun_out/x.json'));print('$1', round(d['value'],4), round(d['ms_per_step'],1), round(d['e2e']['value'],4), round(d['e2e']['ms_per_step'],1), d['kernel_ms'], d['wfa_tier_routing'], d['parity_check']['identical'])"; }
run tier_learn=1
