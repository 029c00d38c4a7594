#!/bin/bash
# round 2, first GPU call: roofline denominators, full-size C3 counters of the voting kernel, a --set full capture of
# the full-size table on a strided subset of reference points
O=gpurun_out; mkdir -p $O
python tools/atoms_microbench.py > $O/r2_atoms_microbench.json 2> $O/r2_atoms_microbench.err; cat $O/r2_atoms_microbench.json
python tools/vote_probe.py --workload c3 --ref-step 50 > $O/r2_probe_c3_base.json 2> $O/r2_probe_c3_base.err; cat $O/r2_probe_c3_base.json
M=gpu__time_duration.sum,smsp__inst_executed_op_shared_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum,lts__t_sectors.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:ppf_vote_kernel -s 3 -c 1 --csv --log-file $O/r2_ncu_c3_full_metrics_base.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > $O/r2_ncu_c3_full_metrics_base.log 2>&1
tail -4 $O/r2_ncu_c3_full_metrics_base.csv | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ppf_vote_kernel -s 1 -c 1 -f -o $O/r2_prof_vote_c3_step50_base \
    python tools/vote_probe.py --workload c3 --ref-step 50 --repeat 2 > $O/r2_ncu_full_c3_step50_base.log 2>&1
tail -3 $O/r2_ncu_full_c3_step50_base.log | cut -c1-300
