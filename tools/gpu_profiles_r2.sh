#!/bin/bash
# round 2: one `ncu --set full` capture per kernel family (reports under gpurun_out/, summaries made on the CPU side with
# tools/summarize_ncu.py), plus the launch list of one default bench step.  usage (GPU box): bash tools/gpu_profiles_r2.sh
O=gpurun_out; mkdir -p $O
cap() {  # cap NAME KERNEL_REGEX SKIP -- command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/r2_prof_$name "$@" > $O/r2_prof_$name.log 2>&1
  tail -1 $O/r2_prof_$name.log | cut -c1-160
  # the reports stay on the GPU box (gpurun brings back at most 64 MiB): only their summaries travel
  python tools/summarize_ncu.py $O/r2_prof_$name.ncu-rep $O/r2_ncu_$name.txt > /dev/null 2>&1
  rm -f $O/r2_prof_$name.ncu-rep
}
SWEEP="python tools/table_build_sweep.py --sizes 10000 --features --cpu-max 0 --repeat 1"
cap k1_features ppf_model_features_kernel 0 $SWEEP
cap k2a_keys keys_from_cloud_kernel 0 $SWEEP
cap radix_scatter radix_scatter_kernel 1 $SWEEP
cap k2c_entries entries_kernel 0 $SWEEP
cap k2c_merge_words merge_words_kernel 0 $SWEEP
cap k2c_bank_order bank_order_kernel 0 $SWEEP
PROBE="python tools/vote_probe.py --workload c3 --ref-step 50 --repeat 2"
cap k3_vote_c3_step50 ppf_vote_kernel 1 $PROBE
cap grid_gather grid_gather_kernel 1 $PROBE
cap k3_vote_c2 ppf_vote_kernel 1 python tools/vote_probe.py --workload c2 --ref-step 1 --repeat 2
cat > /tmp/k4probe.py <<PY
import sys; sys.path.insert(0,".")
from yolo_ppf_pose_estimation_b200 import capi, workloads
wl=workloads.load("c3"); ctx=capi.Context(0)
dm,ds=ctx.upload_cloud(wl.model),ctx.upload_cloud(wl.scene)
t=ctx.table_build_from_cloud(dm,wl.angle_step,wl.dist_step)
hy=ctx.vote(dm,t,ds,0,1)
p,v=ctx.cluster(hy,wl.pos_thr,wl.rot_thr); print("c3 cluster_ms",ctx.timings()["cluster_ms"], v)
PY
cap k4_round0 cluster_round_kernel 0 python /tmp/k4probe.py
cap k4_assign cluster_assign_kernel 0 python /tmp/k4probe.py
cat > /tmp/cvprobe.py <<PY
import sys, numpy as np; sys.path.insert(0,"."); sys.path.insert(0,"tests")
from yolo_ppf_pose_estimation_b200 import capi, synth
from conftest import load_cloud
ctx=capi.Context(0); m=load_cloud("bottle_5mm"); s=synth.synth_scene(100000,2,model_seed=1)
d=capi.CvDetector(ctx,0.025,0.05).train_model(m); print(d.info.n_sampled, ctx.timings())
r,n=d.match(s,1/14.0,0.05); print(n, ctx.timings()["vote_ms"])
PY
cap cv_train cv_train_pairs_kernel 0 python /tmp/cvprobe.py
cap cv_vote cv_vote_kernel 0 python /tmp/cvprobe.py
# launch list of the default bench command (shares of the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_launches_c3.log 2>&1
# DRAM traffic of the full C3 voting launch (metric-only pass)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_op_shared_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum,lts__t_sectors.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:ppf_vote_kernel -s 3 -c 1 --csv --log-file $O/r2_ncu_c3_full_metrics.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > $O/r2_ncu_c3_full_metrics.log 2>&1
ls $O/r2_ncu_*.txt | wc -l
