for v in default pf8 nopf1; do
  if [ $v = default ]; then unset TT_B200_LIB; else export TT_B200_LIB=$PWD/two_tower_recommender_model_b200/lib/variants/$v.so; fi
for cfg in "--L 1 --d 64 --rows 10000000" "--L 20 --d 128 --rows 20000000"; do
  python tools/microbench.py ebc --iters 2 $cfg > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"ebc_backward_update" -c 6 --csv --log-file gpurun_out/ab.csv python tools/microbench.py ebc --iters 3 --warmup 2 $cfg > /dev/null 2>&1
  echo "== $v $cfg"; python tools/ncu_kernel_table.py gpurun_out/ab.csv | cut -c1-125
done; done
