run() { timeout 300 python bench.py --no-cpu --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err; python -c "
import json,sys; d=json.load(open('gpurun_out/b.json')); print(sys.argv[1], 'value Gq/s', round(d['value']/1e9,2), 'serial us', round(d['config']['serial_ms_per_step']*1e3,2), 'graph us', round(d['config']['graph_replay_ms_per_step']*1e3,2), 'frac', round(d['roofline']['frac'],3))" "$1" || tail -3 gpurun_out/b.err; }
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run "mb8"
python tools/stage_times.py 65536 2>&1 | tail -3 | cut -c1-330
python tools/stage_times.py 1048576 2>&1 | grep "kernel span" | tail -1
BB_NVCC_DEFINES="-DBB_MIN_BLOCKS=6" python -m binder_b200.build --force -v 2>&1 | grep -E "registers|spill" | head -2
run "mb6"
python -m binder_b200.build --force > /dev/null 2>&1
