#!/bin/bash
# compare tuning variants on a workload: bash tools/gpu_variants.sh WORKLOAD STEPS lib1 lib2 ... (lib "default" = the built library)
WL=$1; STEPS=$2; shift; shift
for v in "$@"; do
  if [ "$v" = default ]; then unset B200PPF_LIB; else export B200PPF_LIB=$PWD/variants/$v.so; fi
  for e in "" $EXTRA_ENVS; do
    [ -n "$e" ] && export $e
    timeout 300 python bench.py --workload $WL --steps $STEPS --warmup 3 --no-cpu > /tmp/v.json 2>/tmp/v.err || tail -3 /tmp/v.err
    python -c "
import json; d=json.load(open('/tmp/v.json')); print('$WL', '$v', '$e', 'ms', round(d['ms_per_step'],3), 'k3', round(d['roofline']['kernel_ms'],3), 'Gvotes/s', round(d['roofline']['votes_per_sec_in_kernel']/1e9,1), 'slices', d['config']['accumulator_slices'], d['result']['votes'])"
    [ -n "$e" ] && unset ${e%%=*}
  done
done
