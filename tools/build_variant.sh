#!/bin/bash
# tools/build_variant.sh NAME "-DTT_POLY_FWD=1 ..."  -> two_tower_recommender_model_b200/lib/variants/NAME.so
set -e
cd "$(dirname "$0")/.."
P=two_tower_recommender_model_b200
mkdir -p $P/lib/variants /tmp/ttv_$1
for f in $P/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DTT_BUILD $2 -Iinclude -I$P/csrc -c $f -o /tmp/ttv_$1/$(basename $f .cu).o &
done
wait
nvcc -shared -o $P/lib/variants/$1.so /tmp/ttv_$1/*.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo built $P/lib/variants/$1.so
