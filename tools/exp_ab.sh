# same-box A/B: the library built from the working tree vs binder_b200/libbinder_b200_base.so
run() { timeout 300 python bench.py --no-cpu --no-e2e --steps 8000 > gpurun_out/b.json 2> gpurun_out/b.err; python -c "
import json,sys; d=json.load(open('gpurun_out/b.json')); print(sys.argv[1], 'value Gq/s', round(d['value']/1e9,2), 'serial us', round(d['config']['serial_ms_per_step']*1e3,2), 'graph us', round(d['config']['graph_replay_ms_per_step']*1e3,2), 'frac', round(d['roofline']['frac'],3))" "$1" || tail -3 gpurun_out/b.err; }
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run new
cp binder_b200/libbinder_b200.so /tmp/new.so
cp binder_b200/libbinder_b200_base.so binder_b200/libbinder_b200.so
run base
cp /tmp/new.so binder_b200/libbinder_b200.so
run new
python tools/stage_times.py 65536 2>&1 | tail -3 | cut -c1-360
python tools/stage_times.py 1048576 2>&1 | grep "kernel span" | tail -1
