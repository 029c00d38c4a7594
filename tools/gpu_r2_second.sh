#!/bin/bash
# round 2: bin-major accumulator + bank-ordered cells — parity tests, probe timings, conflict counters
O=gpurun_out; mkdir -p $O; TAG=${1:-v7}
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15 | tee $O/r2_pytest_parity_$TAG.log
python tools/vote_probe.py --workload c3 --ref-step 50 --tag $TAG > $O/r2_probe_c3_$TAG.json 2> $O/r2_probe_c3_$TAG.err; cat $O/r2_probe_c3_$TAG.json; tail -2 $O/r2_probe_c3_$TAG.err
B200PPF_NO_BANK_SPREAD=1 python tools/vote_probe.py --workload c3 --ref-step 50 --tag ${TAG}_unordered > $O/r2_probe_c3_${TAG}_unordered.json 2>/dev/null; cat $O/r2_probe_c3_${TAG}_unordered.json
python tools/vote_probe.py --workload c2 --ref-step 1 --tag $TAG > $O/r2_probe_c2_$TAG.json 2>/dev/null; cat $O/r2_probe_c2_$TAG.json
python tools/vote_probe.py --workload c3s --ref-step 1 --tag $TAG > $O/r2_probe_c3s_$TAG.json 2>/dev/null; cat $O/r2_probe_c3s_$TAG.json
M=gpu__time_duration.sum,smsp__inst_executed_op_shared_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum,lts__t_sectors.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum
for v in "" 1; do
  n=$TAG; [ -n "$v" ] && n=${TAG}_unordered
  B200PPF_NO_BANK_SPREAD=$v timeout 600 ncu --metrics $M --clock-control none -k regex:ppf_vote_kernel -s 1 -c 1 --csv --log-file $O/r2_ncu_c3_step50_metrics_$n.csv \
      python tools/vote_probe.py --workload c3 --ref-step 50 --repeat 2 > $O/r2_ncu_c3_step50_metrics_$n.log 2>&1
  python - <<PY
import csv
d={r[-3]:r[-1] for r in csv.reader(open("$O/r2_ncu_c3_step50_metrics_$n.csv")) if len(r)>12 and r[0]!='ID'}
a=float(d['smsp__inst_executed_op_shared_atom.sum']); w=float(d['l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum'])
print("$n", 'ms', float(d['gpu__time_duration.sum'])/1e6, 'atoms wavefronts/inst', w/a, 'l1 %', d['l1tex__throughput.avg.pct_of_peak_sustained_elapsed'], 'issue %', d['smsp__issue_active.avg.pct_of_peak_sustained_active'], 'inst', d['smsp__inst_executed.sum'])
PY
done
for wl in c2 c1; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu > $O/r2_bench_${wl}_$TAG.json 2> $O/r2_bench_${wl}_$TAG.err; python -c "
import json; d=json.load(open('$O/r2_bench_${wl}_$TAG.json')); print('$wl', 'ms/step', round(d['ms_per_step'],3), 'k3_ms', d['roofline']['kernel_ms'], d['result'])"; done
