# one GPU, final code: parity suite, the default bench line and config2's (CPU legs, e2e), the ncu launch list of the default
# command, and `ncu --set full` captures of config3 (service variant, full-size zone) and config2
mkdir -p gpurun_out
TAG=${1:-v}
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
SECONDS=0
timeout 900 python bench.py > gpurun_out/${TAG}_bench_config3.json 2> gpurun_out/${TAG}_bench_config3.err || tail -5 gpurun_out/${TAG}_bench_config3.err
echo "default bench.py wall: ${SECONDS}s"
timeout 600 python bench.py --workload config2 > gpurun_out/${TAG}_bench_config2.json 2> gpurun_out/${TAG}_bench_config2.err || tail -5 gpurun_out/${TAG}_bench_config2.err
SECONDS=0
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err || tail -5 gpurun_out/${TAG}_bench_reference.err
echo "reference arm wall: ${SECONDS}s"
python - <<PY
import json
for w in ('bench_config3','bench_config2','bench_reference'):
    try:
        d=json.load(open('gpurun_out/${TAG}_%s.json'%w)); e=d.get('e2e') or {}; c=d.get('cpu_baseline') or {}
        print(w, 'value %.3f G q/s (depth %s)'%(d['value']/1e9, d['config'].get('batches_in_flight')), 'kern_ms %s'%(d.get('roofline') or {}).get('kernel_ms'), 'frac %s'%(d.get('roofline') or {}).get('frac'), 'traffic %s'%(d.get('roofline') or {}).get('traffic'), 'e2e %.1f M'%(e.get('value',0)/1e6), 'cpu %.2f M'%(c.get('value',0)/1e6), 'same_table', (c.get('same_table') or {}).get('value'), '|', str(d['config'].get('parity'))[:60])
        for k,v in d['config'].get('also_measured',{}).items(): print('   also', k, '%.3f G q/s'%(v['value']/1e9), 'kern_ms %.4f'%v['kernel_ms'], 'frac %.3f'%v['roofline_frac'], v.get('parity','')[:40])
    except Exception as e: print(w, 'ERR', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1; tail -1 gpurun_out/${TAG}_launches_bench.log | cut -c1-120
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resolve_kernel -s 12 -c 1 -o gpurun_out/${TAG}_prof_config3 python bench.py --no-cpu --no-e2e --also none --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu3.log 2>&1; tail -1 gpurun_out/${TAG}_ncu3.log | cut -c1-120
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resolve_kernel -s 12 -c 1 -o gpurun_out/${TAG}_prof_config2 python bench.py --workload config2 --no-cpu --no-e2e --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu2.log 2>&1; tail -1 gpurun_out/${TAG}_ncu2.log | cut -c1-120
