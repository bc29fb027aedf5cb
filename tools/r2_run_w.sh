# one GPU: parity suite + kernel-path numbers (config3 automatic + small, config2) + ncu of config3 at the full-size zone
mkdir -p gpurun_out
TAG=${1:-w}
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/${TAG}_config3.json 2> gpurun_out/${TAG}_config3.err || tail -5 gpurun_out/${TAG}_config3.err
BB_PROFILE=small timeout 600 python bench.py --no-cpu --no-e2e --also none > gpurun_out/${TAG}_config3_small.json 2> gpurun_out/${TAG}_config3_small.err || tail -5 gpurun_out/${TAG}_config3_small.err
timeout 300 python bench.py --workload config2 --no-cpu --no-e2e > gpurun_out/${TAG}_config2.json 2> gpurun_out/${TAG}_config2.err || tail -5 gpurun_out/${TAG}_config2.err
python - <<PY
import json
for w in ('config3','config3_small','config2'):
    try:
        d=json.load(open('gpurun_out/${TAG}_%s.json'%w))
        print(w, 'value %.3f G q/s (depth %s, %s)'%(d['value']/1e9, d['config'].get('batches_in_flight'), {k: round(x*1e3,1) for k,x in d['config']['ms_per_step_by_batches_in_flight'].items()}), 'kern_ms %.4f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])
        for k,v in d['config'].get('also_measured',{}).items(): print('   also', k, '%.3f G q/s'%(v['value']/1e9), 'kern_ms %.4f'%v['kernel_ms'], 'frac %.3f'%v['roofline_frac'])
    except Exception as e: print(w, 'ERR', e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resolve_kernel -s 12 -c 1 -o gpurun_out/${TAG}_prof_config3 python bench.py --no-cpu --no-e2e --also none --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu3.log 2>&1; tail -1 gpurun_out/${TAG}_ncu3.log | cut -c1-120
