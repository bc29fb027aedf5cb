# one GPU: parity suite, the default bench line (config3 + config4, CPU legs, e2e), config2 kernel path, and ncu captures
# (config2 warm-cache stall reasons; config3 under the service variant)
mkdir -p gpurun_out
TAG=${1:-s}
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
SECONDS=0
timeout 900 python bench.py > gpurun_out/${TAG}_bench_config3.json 2> gpurun_out/${TAG}_bench_config3.err || tail -5 gpurun_out/${TAG}_bench_config3.err
echo "default bench.py wall: ${SECONDS}s"
timeout 600 python bench.py --workload config2 --no-cpu --no-e2e > gpurun_out/${TAG}_config2.json 2> gpurun_out/${TAG}_config2.err || tail -5 gpurun_out/${TAG}_config2.err
BB_PROFILE=small timeout 600 python bench.py --no-cpu --no-e2e --also none > gpurun_out/${TAG}_config3_small.json 2> gpurun_out/${TAG}_config3_small.err || tail -5 gpurun_out/${TAG}_config3_small.err
python - <<PY
import json
for w in ('bench_config3','config2','config3_small'):
    try:
        d=json.load(open('gpurun_out/${TAG}_%s.json'%w)); e=d.get('e2e') or {}; c=d.get('cpu_baseline') or {}
        print(w, 'value %.3f G q/s (depth %s, by depth %s)'%(d['value']/1e9, d['config'].get('batches_in_flight'), d['config'].get('ms_per_step_by_batches_in_flight')), 'kern_ms %.4f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.1f M'%(e.get('value',0)/1e6), 'cpu %.2f M'%(c.get('value',0)/1e6), 'same_table', (c.get('same_table') or {}).get('value'), '|', d['config']['parity'][:60])
        for k,v in d['config'].get('also_measured',{}).items(): print('   also', k, '%.3f G q/s'%(v['value']/1e9), 'kern_ms %.4f'%v['kernel_ms'], 'frac %.3f'%v['roofline_frac'], v.get('parity','')[:40])
    except Exception as e: print(w, 'ERR', e)
PY
timeout 300 ncu --set full --cache-control none --clock-control none --import-source on -k regex:resolve_kernel -s 40 -c 1 -o gpurun_out/${TAG}_prof_config2_warm python bench.py --workload config2 --no-cpu --no-e2e --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu2w.log 2>&1; tail -1 gpurun_out/${TAG}_ncu2w.log | cut -c1-200
BB_PROFILE=service timeout 500 ncu --set full --clock-control none --import-source on -k regex:resolve_kernel -s 12 -c 1 -o gpurun_out/${TAG}_prof_config3 python bench.py --no-cpu --no-e2e --also none --steps 4 --warmup 3 --zone-records 3000000 > gpurun_out/${TAG}_ncu3.log 2>&1; tail -1 gpurun_out/${TAG}_ncu3.log | cut -c1-200
