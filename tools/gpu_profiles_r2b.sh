#!/bin/bash
# round 2, after the last K3 changes (heaviest-first order, sub-phase own cells): the K3 captures, the launch list of one
# default bench step and the full-size counter pass again.  usage (GPU box): bash tools/gpu_profiles_r2b.sh
O=gpurun_out; mkdir -p $O
cap() {  # cap NAME KERNEL_REGEX SKIP -- command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/r2_prof_$name "$@" > $O/r2_prof_$name.log 2>&1
  tail -1 $O/r2_prof_$name.log | cut -c1-160
  python tools/summarize_ncu.py $O/r2_prof_$name.ncu-rep $O/r2_ncu_$name.txt > /dev/null 2>&1
  rm -f $O/r2_prof_$name.ncu-rep
}
cap k3_vote_c3_step50 ppf_vote_kernel 1 python tools/vote_probe.py --workload c3 --ref-step 50 --repeat 2
cap k3_vote_c2 ppf_vote_kernel 1 python tools/vote_probe.py --workload c2 --ref-step 1 --repeat 2
cap k3_ref_cost ref_cost_kernel 1 python tools/vote_probe.py --workload c3 --ref-step 50 --repeat 2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_launches_c3.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_op_shared_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum,lts__t_sectors.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:ppf_vote_kernel -s 3 -c 1 --csv --log-file $O/r2_ncu_c3_full_metrics.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > $O/r2_ncu_c3_full_metrics.log 2>&1
tail -3 $O/r2_ncu_c3_full_metrics.csv | cut -c1-300
