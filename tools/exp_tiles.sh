run() { timeout 300 python bench.py --no-cpu --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err; python -c "
import json,sys; d=json.load(open('gpurun_out/b.json')); print(sys.argv[1], 'value Gq/s', round(d['value']/1e9,2), 'serial us', round(d['config']['serial_ms_per_step']*1e3,2), 'graph us', round(d['config']['graph_replay_ms_per_step']*1e3,2))" "$1"; }
python -m binder_b200.build --force > /dev/null 2>&1; BB_IN_FLIGHT=4 run "T128 if4"; BB_IN_FLIGHT=8 run "T128 if8"; BB_IN_FLIGHT=12 run "T128 if12"
BB_NVCC_DEFINES="-DBB_T=64 -DBB_MIN_BLOCKS=16" python -m binder_b200.build --force -v 2>&1 | grep -E "registers" | head -1
BB_IN_FLIGHT=4 run "T64 if4"; BB_IN_FLIGHT=8 run "T64 if8"
python tools/stage_times.py 65536 2>&1 | tail -3 | cut -c1-330
python -m binder_b200.build --force > /dev/null 2>&1
