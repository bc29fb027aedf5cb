mkdir -p gpurun_out
for g in 0 32 64 128; do
  echo "== granularity $g"; ./tools/micro/probe_gran $g
  ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gran_$g.csv ./tools/micro/probe_gran $g > /dev/null 2>&1
done
