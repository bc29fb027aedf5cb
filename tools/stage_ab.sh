# per-stage tile timings of config3 for compile-time variants (each run builds its own library with the stamps compiled in)
mkdir -p gpurun_out
for v in "" "-DBB_DIRECT_BIG" "-DBB_NO_STREAM_HINT" "-DBB_DIRECT_BIG -DBB_NO_STREAM_HINT"; do
  echo "=== variant: [$v]"
  BB_NVCC_DEFINES="$v" BB_WL=config3 BB_ZONE=3000000 timeout 300 python tools/stage_times.py 262144 2>&1 | grep -v Warning | tail -4 | cut -c1-600
done > gpurun_out/stage_ab.txt 2>&1
python -m binder_b200.build --force > /dev/null 2>&1
cat gpurun_out/stage_ab.txt
