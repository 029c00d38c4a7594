#!/bin/bash
# build a tuning variant of libb200ppf.so into variants/NAME.so:  tools/build_variant.sh NAME [-DFLAG=...]
#   SRC_ROOT=<checkout> builds the sources of another checkout (e.g. a git worktree of an older commit)
# run it with B200PPF_LIB=variants/NAME.so python bench.py ...
NAME=$1; shift
ROOT=${SRC_ROOT:-.}
D=$(mktemp -d)
SRC=$ROOT/yolo_ppf_pose_estimation_b200/csrc
for f in $SRC/*.cu; do
  b=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
       -Xcompiler -fPIC -Xcompiler -O2 "$@" -c $f -o $D/$b.o &
done
wait
mkdir -p variants
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/$NAME.so $D/*.o -cudart shared && echo built variants/$NAME.so
rm -rf $D
