#!/bin/bash
# scaling of the default bench on N GPUs of one box: bash tools/gpu_scale.sh N TAG
N=$1; TAG=${2:-r2}; O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > $O/r2_bench_c3_n${N}_$TAG.json 2> $O/r2_bench_c3_n${N}_$TAG.err
tail -c 400 $O/r2_bench_c3_n${N}_$TAG.err
python -c "
import json; d=[json.loads(l) for l in open('$O/r2_bench_c3_n${N}_$TAG.json') if l.startswith('{')][-1]; r=d['roofline']; print('N=$N ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'k3 mean', round(r['kernel_ms'],2), 'k3 max', round(r['kernel_ms_max_over_ranks'],2), d['result']['votes'])"
