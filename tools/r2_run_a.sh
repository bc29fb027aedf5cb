# round 2, first GPU call: the new bench on the round-1 kernels (baseline for every later change)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2a_bench_n1_config3.json 2> gpurun_out/r2a_bench_n1_config3.err; tail -c 1500 gpurun_out/r2a_bench_n1_config3.json; tail -5 gpurun_out/r2a_bench_n1_config3.err
timeout 300 python bench.py --workload config2 > gpurun_out/r2a_bench_n1_config2.json 2> gpurun_out/r2a_bench_n1_config2.err; tail -c 800 gpurun_out/r2a_bench_n1_config2.json; tail -3 gpurun_out/r2a_bench_n1_config2.err
timeout 500 ncu --set full --clock-control none --import-source on -k regex:resolve_kernel -s 12 -c 1 -o gpurun_out/r2a_prof_config3 python bench.py --no-cpu --no-e2e --also none --steps 4 --warmup 3 > gpurun_out/r2a_ncu.log 2>&1; tail -2 gpurun_out/r2a_ncu.log
