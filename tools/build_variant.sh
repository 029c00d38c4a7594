#!/bin/bash
# build a tuning variant of libb200ppf.so into variants/NAME.so:  tools/build_variant.sh NAME -DFLAG=... 
# run it with B200PPF_LIB=variants/NAME.so python bench.py ...
NAME=$1; shift
D=$(mktemp -d)
SRC=yolo_ppf_pose_estimation_b200/csrc
for f in capi radix_sort k1_features k2_table k3_vote scene_grid k4_cluster k5_transform k6_icp microbench; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
       -Xcompiler -fPIC -Xcompiler -O2 "$@" -c $SRC/$f.cu -o $D/$f.o &
done
wait
mkdir -p variants
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/$NAME.so $D/*.o -cudart shared && echo built variants/$NAME.so
rm -rf $D
