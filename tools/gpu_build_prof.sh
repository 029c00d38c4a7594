#!/bin/bash
# per-kernel times of one table build (ncu launch list) + the C5 sweep.  usage: bash tools/gpu_build_prof.sh TAG [sizes]
TAG=${1:-b}; SIZES=${2:-5000,10000,20000}
O=gpurun_out; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/build_launches_$TAG.csv \
    python tools/table_build_sweep.py --sizes 10000 --repeat 1 > $O/build_ncu_$TAG.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("$O/build_launches_$TAG.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > mv:
        try: agg[r[kn]] = agg.get(r[kn], 0.0) + float(r[mv].replace(",", ""))
        except ValueError: pass
for k, v in agg.items(): print(f"{v/1e6:9.3f} ms  {k[:110]}")
PY
python tools/table_build_sweep.py --sizes $SIZES | tee $O/table_sweep_$TAG.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in d.items() if k in ('n_model','keys_ms','sort_ms','csr_ms','device_ms','pairs_per_sec','slices','key_bits')})"
