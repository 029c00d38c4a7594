#!/bin/bash
# One gpurun call: GPU parity tests, bench lines (C3 default with the CPU leg and the reference arm, C2, C1, C2-5mm, C4),
# ncu launch list of the default command, full captures of the voting kernel (C2 and quarter-scale C3), ICP timing.  usage (repo root on the GPU box): bash tools/gpu_round.sh TAG
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest_$TAG.log
grep -q failed $O/pytest_$TAG.log && exit 1
python bench.py > $O/bench_c3_$TAG.json 2> $O/bench_c3_$TAG.err; tail -c 600 $O/bench_c3_$TAG.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_c3_ref_$TAG.json 2> $O/bench_c3_ref_$TAG.err
python bench.py --workload c2 --steps 10 --warmup 3 > $O/bench_c2_$TAG.json 2> $O/bench_c2_$TAG.err; tail -c 600 $O/bench_c2_$TAG.err
timeout 300 python bench.py --workload c2 --impl reference --steps 3 --warmup 1 > $O/bench_c2_ref_$TAG.json 2> $O/bench_c2_ref_$TAG.err
timeout 300 python bench.py --workload c1 --steps 10 --warmup 3 > $O/bench_c1_$TAG.json 2> $O/bench_c1_$TAG.err
timeout 300 python bench.py --workload c2_5mm --steps 3 --warmup 3 --no-cpu > $O/bench_c2_5mm_$TAG.json 2> $O/bench_c2_5mm_$TAG.err
timeout 300 python bench.py --workload c4 --steps 2 --warmup 3 > $O/bench_c4_$TAG.json 2> $O/bench_c4_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_launches_$TAG.log 2>&1
bash tools/gpu_prof.sh $TAG c2
bash tools/gpu_prof.sh $TAG c3s
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ppf_vote_kernel -s 3 -c 1 --csv \
    --log-file $O/dram_c3_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_dram_c3_$TAG.log 2>&1
python tools/icp_bench.py > $O/icp_bench_$TAG.json 2> $O/icp_bench_$TAG.err
# pre-processing row: stage and per-object timings vs the CPU chain, launch list, full captures of the neighbour kernel
timeout 300 python tools/prep_bench.py > $O/prep_bench_$TAG.jsonl 2> $O/prep_bench_$TAG.err; tail -c 400 $O/prep_bench_$TAG.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/prep_launches_$TAG.csv \
    python tools/prep_bench.py --no-cpu --repeat 1 --only big --no-frames > $O/ncu_prep_launches_$TAG.log 2>&1
for which in sor:2 normals:3; do
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:knn_ -s ${which##*:} -c 1 -f -o $O/prof_knn_${which%%:*}_$TAG \
        python tools/prep_bench.py --no-cpu --repeat 1 --only big --no-frames > $O/ncu_full_knn_${which%%:*}_$TAG.log 2>&1
done
for f in c3 c3_ref c2 c2_ref c1 c2_5mm c4; do python - <<PY
import json
try:
    d = json.load(open("$O/bench_${f}_$TAG.json"))
    r = d.get("roofline", {})
    print("$f", "value", "%.4g" % d["value"], "ms/step", round(d["ms_per_step"], 3), "e2e", "%.4g" % d["e2e"]["value"],
          "k3_ms", r.get("kernel_ms"), "atomic_frac", (r.get("atomic") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$f", "no line:", e)
PY
done
