show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], 'value Gq/s', round(d['value']/1e9,2), 'serial us', round(d['config']['serial_ms_per_step']*1e3,1), 'e2e Mq/s', round(d['e2e']['value']/1e6,1) if d['e2e'] else None, 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']/1e6,2))" "$1" "$2" || tail -3 gpurun_out/w.err; }
timeout 500 python bench.py --workload config3 --zone-records 3000000 --batch 262144 --steps 60 > gpurun_out/bench_r1_f_config3.json 2> gpurun_out/w.err; show gpurun_out/bench_r1_f_config3.json config3
timeout 300 python bench.py --workload config5 --steps 200 > gpurun_out/bench_r1_f_config5.json 2> gpurun_out/w.err; show gpurun_out/bench_r1_f_config5.json config5
timeout 200 python -m pytest tests/test_balancer_frames.py tests/test_config1_udp.py tests/test_wire_golden.py -q -m gpu 2>&1 | tail -2
