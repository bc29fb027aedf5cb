# A/B on one GPU: library variants x {config3 (automatic = service variant), config3 small variant, config2}, kernel path only
mkdir -p gpurun_out
cp binder_b200/libbinder_b200.so /tmp/lib_default.so
for v in ${VARIANTS:-base noblk win16 base}; do
  cp binder_b200/variants/lib_$v.so binder_b200/libbinder_b200.so
  timeout 300 python bench.py --no-cpu --no-e2e --also none --zone-records 3000000 > gpurun_out/ab4_${v}_c3.json 2> gpurun_out/ab4_${v}.err
  BB_PROFILE=small timeout 300 python bench.py --no-cpu --no-e2e --also none --zone-records 3000000 > gpurun_out/ab4_${v}_c3small.json 2>> gpurun_out/ab4_${v}.err
  timeout 200 python bench.py --workload config2 --no-cpu --no-e2e > gpurun_out/ab4_${v}_c2.json 2>> gpurun_out/ab4_${v}.err
  python - <<PY
import json
out=['$v:']
for w in ('c3','c3small','c2'):
    try:
        d=json.load(open('gpurun_out/ab4_${v}_%s.json'%w)); out.append('%s %.3f G (kern %.1f us, depths %s)'%(w, d['value']/1e9, d['roofline']['kernel_ms']*1e3, {k: round(x*1e3,1) for k,x in d['config']['ms_per_step_by_batches_in_flight'].items()}))
    except Exception as e: out.append('%s ERR %r'%(w,e))
print('  '.join(out))
PY
done
cp /tmp/lib_default.so binder_b200/libbinder_b200.so
