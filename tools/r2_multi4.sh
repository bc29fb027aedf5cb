# 4 GPUs: config 4 with the zone sharded (peer stores), every rank checked against the oracle; then the world-4 parity tests
mkdir -p gpurun_out
TAG=${1:-m4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29662 bench.py --gpus 4 --mode shard > gpurun_out/${TAG}_n4_shard.json 2> gpurun_out/${TAG}_n4_shard.err || tail -8 gpurun_out/${TAG}_n4_shard.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_n4_shard.json')); e=d.get('e2e') or {}
    print('N=4 shard: value %.3f G q/s  ms/step %.4f  e2e %.1f M q/s  | %s'%(d['value']/1e9, d['ms_per_step'], e.get('value',0)/1e6, d['config']['parity'][:80]))
except Exception as ex: print('N=4 shard ERR', ex)
PY
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -k "0-0-4 or 0-1-4" > gpurun_out/${TAG}_pytest_multi4.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_multi4.log
