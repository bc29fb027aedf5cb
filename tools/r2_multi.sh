# multi-GPU: parity tests for the worlds this box has, then config 4 in the three modes (each with the oracle check of every rank)
N=${1:-2}; TAG=${2:-m}
mkdir -p gpurun_out
[ "${4:-0}" = "1" ] || timeout 1500 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/${TAG}_pytest_multi.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_multi.log
MODES=${3:-"shard replicas nccl"}
SKIPTESTS=${4:-0}
for mode in $MODES; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --mode $mode > gpurun_out/${TAG}_n${N}_${mode}.json 2> gpurun_out/${TAG}_n${N}_${mode}.err || tail -8 gpurun_out/${TAG}_n${N}_${mode}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_n${N}_${mode}.json')); e=d.get('e2e') or {}
    print('N=$N $mode: value %.3f G q/s  ms/step %.4f  e2e %.1f M q/s  | %s'%(d['value']/1e9, d['ms_per_step'], e.get('value',0)/1e6, d['config']['parity'][:70]))
except Exception as ex: print('N=$N $mode ERR', ex)
PY
done
