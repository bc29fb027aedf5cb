# round-end validation on one GPU: parity suite, smoke, default bench line, ncu launch list + full capture
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final_n1.json; tail -2 gpurun_out/bench_final.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_l.log 2>&1; tail -1 gpurun_out/ncu_l.log | cut -c1-200
timeout 200 ncu --set full --clock-control none --import-source on -k regex:resolve_kernel -c 1 -s 30 -o gpurun_out/prof_final python bench.py --no-cpu --no-e2e --steps 40 --warmup 3 > gpurun_out/ncu_f.log 2>&1; tail -1 gpurun_out/ncu_f.log
