# 2 GPUs: config 5 (90 % misses, recursion split) on more than one GPU, default mode, every rank checked against the oracle;
# plus the tightened balancer-frames GPU test
mkdir -p gpurun_out
TAG=${1:-m2c}
timeout 300 python -m pytest tests/test_balancer_frames.py -x -q -m gpu > gpurun_out/${TAG}_pytest_frames.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_frames.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --workload config5 > gpurun_out/${TAG}_n2_config5.json 2> gpurun_out/${TAG}_n2_config5.err || tail -8 gpurun_out/${TAG}_n2_config5.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_n2_config5.json')); e=d.get('e2e') or {}
    print('N=2 config5: mode %s value %.3f G q/s  ms/step %.4f  e2e %.1f M q/s  | %s'%(d['config']['mode'], d['value']/1e9, d['ms_per_step'], e.get('value',0)/1e6, d['config']['parity'][:90]))
except Exception as ex: print('N=2 config5 ERR', ex)
PY
