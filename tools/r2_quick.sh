# quick A/B on one GPU: parity subset + kernel-path numbers of config3(+config4) and config2 (no oracle, no e2e)
mkdir -p gpurun_out
TAG=${1:-q}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/${TAG}_config3.json 2> gpurun_out/${TAG}_config3.err || tail -5 gpurun_out/${TAG}_config3.err
timeout 300 python bench.py --workload config2 --no-cpu --no-e2e > gpurun_out/${TAG}_config2.json 2> gpurun_out/${TAG}_config2.err || tail -5 gpurun_out/${TAG}_config2.err
python - <<PY
import json
for w in ('config3','config2'):
    try:
        d=json.load(open('gpurun_out/${TAG}_%s.json'%w))
        print(w, 'value %.3f G q/s'%(d['value']/1e9), 'kern_ms %.4f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], 'inflight %.3f'%d['roofline']['in_flight_frac'])
        for k,v in d['config'].get('also_measured',{}).items(): print('   also', k, '%.3f G q/s'%(v['value']/1e9), 'frac %.3f'%v['roofline_frac'])
    except Exception as e: print(w, 'ERR', e)
PY
