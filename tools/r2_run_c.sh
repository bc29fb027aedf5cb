# round 2, call c: + sequential probe, 32-byte L2 fetches, instruction diet: parity suite, bench config3 (+config4) and config2, ncu of both kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/r2c_bench_n1_config3.json 2> gpurun_out/r2c_bench_n1_config3.err; tail -c 600 gpurun_out/r2c_bench_n1_config3.json; tail -3 gpurun_out/r2c_bench_n1_config3.err
timeout 300 python bench.py --workload config2 > gpurun_out/r2c_bench_n1_config2.json 2> gpurun_out/r2c_bench_n1_config2.err; tail -c 400 gpurun_out/r2c_bench_n1_config2.json; tail -3 gpurun_out/r2c_bench_n1_config2.err
timeout 500 ncu --set full --clock-control none --import-source on -k regex:resolve_kernel -s 12 -c 1 -o gpurun_out/r2c_prof_config3 python bench.py --no-cpu --no-e2e --also none --steps 4 --warmup 3 > gpurun_out/r2c_ncu3.log 2>&1; tail -2 gpurun_out/r2c_ncu3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resolve_kernel -s 12 -c 1 -o gpurun_out/r2c_prof_config2 python bench.py --workload config2 --no-cpu --no-e2e --steps 4 --warmup 3 > gpurun_out/r2c_ncu2.log 2>&1; tail -2 gpurun_out/r2c_ncu2.log
