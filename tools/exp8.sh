mkdir -p gpurun_out
for v in p0 p1 p2 p3; do
  AUGB200_LIB=$PWD/build_variants/$v.so timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_$v.csv python tools/prof_sweep.py 1184 1 > /dev/null 2>&1
  echo $v $(grep k_prep gpurun_out/l_$v.csv | tail -1 | awk -F'","' '{print $NF}')
done
