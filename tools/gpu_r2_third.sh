#!/bin/bash
O=gpurun_out; mkdir -p $O; TAG=${1:-v8}
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/r2_pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > $O/r2_bench_c3_$TAG.json 2> $O/r2_bench_c3_$TAG.err; tail -c 300 $O/r2_bench_c3_$TAG.err
python -c "
import json; d=json.load(open('$O/r2_bench_c3_$TAG.json')); r=d['roofline']; print('c3 ms/step', d['ms_per_step'], 'k3', r['kernel_ms'], 'frac', r['frac'], 'votes/s', r['achieved'], 'peak', r['peak'], 'stage', d['table_build_stage_ms'], d['table_build_ms'], d['result'])"
