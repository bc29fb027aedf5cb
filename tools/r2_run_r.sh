# one GPU: parity suite, then the three full bench lines (default = config3 + config4; config2; config5 + its RD=0 twin) with the
# CPU legs and e2e, then a warm-cache ncu pass of config2 (does the instruction-fetch stall share survive warm caches?)
mkdir -p gpurun_out
TAG=${1:-r}
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
/usr/bin/time -v -o gpurun_out/${TAG}_time_default.txt timeout 900 python bench.py > gpurun_out/${TAG}_bench_config3.json 2> gpurun_out/${TAG}_bench_config3.err || tail -5 gpurun_out/${TAG}_bench_config3.err
timeout 600 python bench.py --workload config2 > gpurun_out/${TAG}_bench_config2.json 2> gpurun_out/${TAG}_bench_config2.err || tail -5 gpurun_out/${TAG}_bench_config2.err
timeout 900 python bench.py --workload config5 > gpurun_out/${TAG}_bench_config5.json 2> gpurun_out/${TAG}_bench_config5.err || tail -5 gpurun_out/${TAG}_bench_config5.err
python - <<PY
import json
for w in ('config3','config2','config5'):
    try:
        d=json.load(open('gpurun_out/${TAG}_bench_%s.json'%w)); e=d.get('e2e') or {}; c=d.get('cpu_baseline') or {}
        print(w, 'value %.3f G q/s (depth %s)'%(d['value']/1e9, d['config'].get('batches_in_flight')), 'kern_ms %.4f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.1f M'%(e.get('value',0)/1e6), 'cpu %.2f M'%(c.get('value',0)/1e6), 'same_table', (c.get('same_table') or {}).get('value'), '|', d['config']['parity'][:60])
        for k,v in d['config'].get('also_measured',{}).items(): print('   also', k, '%.3f G q/s'%(v['value']/1e9), 'kern_ms %.4f'%v['kernel_ms'], 'frac %.3f'%v['roofline_frac'], v.get('parity','')[:40])
    except Exception as e: print(w, 'ERR', e)
PY
grep -E "Elapsed|Maximum resident" gpurun_out/${TAG}_time_default.txt
timeout 300 ncu --set full --cache-control none --clock-control none --import-source on -k regex:resolve_kernel -s 40 -c 1 -o gpurun_out/${TAG}_prof_config2_warm python bench.py --workload config2 --no-cpu --no-e2e --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu2w.log 2>&1; tail -1 gpurun_out/${TAG}_ncu2w.log | cut -c1-200
