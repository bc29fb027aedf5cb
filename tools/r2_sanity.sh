# what the driver does at round end, on one GPU: smoke(), the -m gpu tests, the default bench line and the reference arm
mkdir -p gpurun_out
TAG=${1:-z}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
SECONDS=0
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -5 gpurun_out/${TAG}_bench.err
echo "bench.py --steps 20 --warmup 3 wall: ${SECONDS}s"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); e=d.get('e2e') or {}; c=d.get('cpu_baseline') or {}; r=d['roofline']
print('value %.3f G q/s (depth %s) kern_ms %.4f frac %.3f traffic %s e2e %.1f M cpu %.2f M same_table %s | %s' % (d['value']/1e9, d['config'].get('batches_in_flight'), r['kernel_ms'], r['frac'], r['traffic'], e.get('value',0)/1e6, c.get('value',0)/1e6, (c.get('same_table') or {}).get('value'), d['config']['parity'][:50]))
PY
