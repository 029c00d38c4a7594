#!/bin/bash
# round 2 closing run on one GPU: the GPU test suite, smoke, the default bench line with its CPU leg, the reference arm, the other workloads
O=gpurun_out; mkdir -p $O; TAG=${1:-final}
timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12 | tee $O/r2_pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py --steps 10 --warmup 3 > $O/r2_bench_c3_$TAG.json 2> $O/r2_bench_c3_$TAG.err; tail -c 300 $O/r2_bench_c3_$TAG.err
timeout 1500 python bench.py --impl reference --steps 10 --warmup 2 > $O/r2_bench_c3_ref_$TAG.json 2> $O/r2_bench_c3_ref_$TAG.err; tail -c 300 $O/r2_bench_c3_ref_$TAG.err
for wl in c2 c1 c2_5mm; do timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > $O/r2_bench_${wl}_$TAG.json 2> $O/r2_bench_${wl}_$TAG.err; done
timeout 600 python bench.py --workload c2 --impl reference --steps 5 --warmup 1 > $O/r2_bench_c2_ref_$TAG.json 2> $O/r2_bench_c2_ref_$TAG.err
timeout 900 python bench.py --workload c4 --steps 2 --warmup 3 > $O/r2_bench_c4_$TAG.json 2> $O/r2_bench_c4_$TAG.err
for f in c3 c3_ref c2 c2_ref c1 c2_5mm c4; do python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$O/r2_bench_${f}_$TAG.json") if l.startswith("{")][-1]
    r=d.get("roofline",{})
    print("$f", "value %.4g" % d["value"], "ms/step", round(d["ms_per_step"],3), "e2e %.4g" % d["e2e"]["value"], "k3_ms", r.get("kernel_ms"), "frac", r.get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"), "votes/s %.4g" % d.get("votes_per_sec",0))
except Exception as e:
    print("$f", "no line:", e)
PY
done
