# config3/config4/config2 kernel-path numbers for: default, no stream hints, 128-byte L2 fetches, 128-byte + no hints
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --no-cpu --no-e2e > gpurun_out/ab3_$tag.json 2> gpurun_out/ab3_$tag.err; env "$@" timeout 200 python bench.py --workload config2 --no-cpu --no-e2e > gpurun_out/ab3_${tag}_c2.json 2>> gpurun_out/ab3_$tag.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ab3_$tag.json')); a=d['config']['also_measured'].get('config4',{}); c=json.load(open('gpurun_out/ab3_${tag}_c2.json'))
    print('$tag: config3 %.3f G (kern %.1f us)  config4 %.3f G  config2 %.3f G (kern %.2f us)'%(d['value']/1e9, d['roofline']['kernel_ms']*1e3, a.get('value',0)/1e9, c['value']/1e9, c['roofline']['kernel_ms']*1e3))
except Exception as e: print('$tag ERR', e)
PY
}
run default BB_X=1
run gran128 BB_L2_GRAN=128
BB_NVCC_DEFINES="-DBB_NO_STREAM_HINT" python -m binder_b200.build --force > /dev/null 2>&1
run nohint BB_X=1
run nohint_gran128 BB_L2_GRAN=128
python -m binder_b200.build --force > /dev/null 2>&1
