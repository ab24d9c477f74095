#!/bin/bash
# Round-2 ncu captures (run under gpurun on one B200): launch list of the bench command, --set full of the dominant
# kernels.  The .ncu-rep files stay on the box (gpurun_out/ is capped at 64 MiB); their details / raw pages come back as CSV.
set -x
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches_bench_py.csv python bench.py --steps 4 --warmup 3 --no-configs --no-cpu-baseline > $O/r2_bench_under_ncu.json 2> /dev/null
cap() {   # kernel regex, skip, driver...
  k=$1; skip=$2; shift 2
  ncu --set full --import-source on --clock-control none -k regex:$k -s $skip -c 1 -f -o /tmp/r2_full_$k "$@" > /dev/null 2>&1
  ncu -i /tmp/r2_full_$k.ncu-rep --page details --csv > $O/r2_ncu_${k}_details.csv 2>/dev/null
  ncu -i /tmp/r2_full_$k.ncu-rep --page raw --csv > $O/r2_ncu_${k}_raw.csv 2>/dev/null
  rm -f /tmp/r2_full_$k.ncu-rep
}
for k in k_accumulate k_prepare k_scalars k_scatter; do cap $k 1 python tools/dev_verify_once.py 20 2; done
for k in k_dec_finish; do cap $k 0 python tools/dev_wire_once.py 18; done
for k in k_ell2_maps k_scalar_mul_plan k_scalar_mul_proj; do cap $k 0 python tools/dev_feeders_once.py 20; done
ls -la $O/r2_ncu_* $O/r2_launches_bench_py.csv
