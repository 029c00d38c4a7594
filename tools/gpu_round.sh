#!/bin/bash
# One gpurun call: GPU parity tests, bench (C2 + probes), microbench, ncu launch list and one full capture.
# usage (from the repo root on the GPU box): bash tools/gpu_round.sh TAG
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest_$TAG.log
python bench.py --steps 10 --warmup 3 > $O/bench_c2_$TAG.json 2> $O/bench_c2_$TAG.err
tail -c 600 $O/bench_c2_$TAG.err
python - <<'PY' 2>&1 | tee gpurun_out/atoms_$TAG.json
import json
from yolo_ppf_pose_estimation_b200 import capi
c = capi.Context(0)
print(json.dumps({f"pattern{p}": c.microbench_atoms(p) for p in (0, 1, 2)}))
PY
timeout 600 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > $O/bench_c3_$TAG.json 2> $O/bench_c3_$TAG.err
tail -c 600 $O/bench_c3_$TAG.err
timeout 300 python bench.py --workload c2_5mm --steps 3 --warmup 3 --no-cpu > $O/bench_c2_5mm_$TAG.json 2> $O/bench_c2_5mm_$TAG.err
timeout 300 python bench.py --workload c1 --steps 10 --warmup 3 > $O/bench_c1_$TAG.json 2> $O/bench_c1_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ppf_vote_kernel -s 3 -c 1 -f -o $O/prof_vote_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_full_$TAG.log 2>&1
ls -la $O | tail -20
