#!/bin/bash
# One gpurun call for the pre-processing row: GPU tests, timing vs the CPU restatement, ncu launch list and full
# captures of the neighbour kernel (outlier-removal and normal epilogues) on the 1 Mi-point scene.
# usage (repo root on the GPU box): bash tools/gpu_prep_prof.sh TAG [full]   ("full" also runs the whole GPU suite)
TAG=${1:-p1}
O=gpurun_out
mkdir -p $O
if [ "$2" = "full" ]; then
    python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest_$TAG.log
else
    python -m pytest tests/test_gpu_prep.py -x -q 2>&1 | tail -5 | tee $O/pytest_$TAG.log
fi
timeout 300 python tools/prep_bench.py > $O/prep_bench_$TAG.json 2> $O/prep_bench_$TAG.err; cat $O/prep_bench_$TAG.json; tail -3 $O/prep_bench_$TAG.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/prep_launches_$TAG.csv \
    python tools/prep_bench.py --no-cpu --repeat 1 --only big --no-frames > $O/ncu_prep_launches_$TAG.log 2>&1
for which in sor:2 normals:3; do
    name=${which%%:*}; skip=${which##*:}
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:knn_ -s $skip -c 1 -f -o $O/prof_knn_${name}_$TAG \
        python tools/prep_bench.py --no-cpu --repeat 1 --only big --no-frames > $O/ncu_full_knn_${name}_$TAG.log 2>&1
    tail -2 $O/ncu_full_knn_${name}_$TAG.log | cut -c1-200
done
