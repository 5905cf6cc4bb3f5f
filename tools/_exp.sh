run() { MGB_PARAMS=$1 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/x.json 2> gpurun_out/x.err; tail -2 gpurun_out/x.err; python -c "
import json;d=json.load(open('gpurun_out/x.json'));print('$1', round(d['value'],4), round(d['ms_per_step'],1), d['kernel_ms']['k_wfa_big'], d['device_cycles_last_step']['wfa_max_cyc'])"; }
run mb7=8
run mb7=6
run mb7=4
