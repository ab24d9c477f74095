#!/bin/bash
# Round-2 ncu captures (run under gpurun on one B200): launch list of the bench command, --set full of the dominant kernels.
set -x
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches_bench_py.csv python bench.py --steps 4 --warmup 3 --no-configs --no-cpu-baseline > $O/r2_bench_under_ncu.json 2> /dev/null
for k in k_accumulate k_prepare k_scalars k_scatter; do
  ncu --set full --import-source on --clock-control none -k regex:$k -s 1 -c 1 -f -o $O/r2_full_$k python tools/dev_verify_once.py 20 2 > /dev/null 2>&1
done
for k in k_ell2_maps k_scalar_mul_proj k_dec_finish; do
  ncu --set full --import-source on --clock-control none -k regex:$k -c 2 -f -o $O/r2_full_$k python tools/dev_feeders_once.py 20 > /dev/null 2>&1
done
ls -la $O/*.ncu-rep
