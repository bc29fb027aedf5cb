# 2 GPUs: the sharded parity tests (release/acquire epoch flags), where a sharded step's time goes + measured NVLink bytes,
# then config 4 as the default (auto = replicas) with every rank's answers checked against the oracle
mkdir -p gpurun_out
TAG=${1:-m2}
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/${TAG}_pytest_multi.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tools/multi_phase_times.py > gpurun_out/${TAG}_phase_times.txt 2> gpurun_out/${TAG}_phase_times.err; cat gpurun_out/${TAG}_phase_times.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 > gpurun_out/${TAG}_n2_auto.json 2> gpurun_out/${TAG}_n2_auto.err || tail -8 gpurun_out/${TAG}_n2_auto.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_n2_auto.json')); e=d.get('e2e') or {}
    print('N=2 auto: mode %s value %.3f G q/s  ms/step %.4f  e2e %.1f M q/s  | %s'%(d['config']['mode'], d['value']/1e9, d['ms_per_step'], e.get('value',0)/1e6, d['config']['parity'][:70]))
except Exception as ex: print('N=2 auto ERR', ex)
PY
