#!/bin/bash
# quick GPU check: parity tests, then C2 / C3 bench lines.  usage: bash tools/gpu_quick.sh TAG [c3steps]
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_$TAG.log
grep -q failed $O/pytest_$TAG.log && exit 1
python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_c2_$TAG.json 2> $O/bench_c2_$TAG.err; tail -c 400 $O/bench_c2_$TAG.err
python -c "
import json; d=json.load(open('$O/bench_c2_$TAG.json')); print('C2 ms', d['ms_per_step'], 'k3', d['roofline']['kernel_ms'], 'votes/s', d['roofline']['votes_per_sec_in_kernel'], d['result'])"
if [ -n "$2" ]; then
timeout 240 python bench.py --workload c3 --steps $2 --warmup 3 --no-cpu > $O/bench_c3_$TAG.json 2> $O/bench_c3_$TAG.err; tail -c 400 $O/bench_c3_$TAG.err
python -c "
import json; d=json.load(open('$O/bench_c3_$TAG.json')); print('C3 ms', d['ms_per_step'], 'k3', d['roofline']['kernel_ms'], 'votes/s', d['roofline']['votes_per_sec_in_kernel'], d['result'], d['config']['accumulator_slices'])"
fi
