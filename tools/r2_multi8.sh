# 8 GPUs: the world-8 parity tests (peer stores and the collective exchange), then config 4 sharded and as replicas
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -k "0-0-8 or 0-1-8 or (collective and 8)" > gpurun_out/r2o_pytest_multi8.log 2>&1; tail -3 gpurun_out/r2o_pytest_multi8.log
bash tools/r2_multi.sh 8 r2o "shard replicas" 1
