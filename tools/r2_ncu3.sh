mkdir -p gpurun_out
TAG=${1:-x}
timeout 500 ncu --set full --clock-control none --import-source on -k regex:resolve_kernel -s 12 -c 1 -o gpurun_out/${TAG}_prof_config3 python bench.py --no-cpu --no-e2e --also none --steps 4 --warmup 3 --zone-records 3000000 > gpurun_out/${TAG}_ncu3.log 2>&1; tail -2 gpurun_out/${TAG}_ncu3.log
