for v in default minb8; do
  if [ $v = default ]; then unset TT_B200_LIB; else export TT_B200_LIB=$PWD/two_tower_recommender_model_b200/lib/variants/$v.so; fi
  echo "== $v"; python tools/run_configs.py 3 | cut -c100-330
  for cfg in "--L 1 --d 64 --rows 10000000" "--L 20 --d 128 --rows 20000000"; do
    python tools/microbench.py ebc --iters 2 $cfg > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ebc_backward_update" -c 4 --csv --log-file gpurun_out/ab.csv python tools/microbench.py ebc --iters 3 --warmup 1 $cfg > /dev/null 2>&1
    echo "   $cfg"; python tools/ncu_kernel_table.py gpurun_out/ab.csv | cut -c1-30
  done
done
