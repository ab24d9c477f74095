#!/bin/bash
# Round-2 final single-GPU record: GPU tests, the driver's bench command for both arms, K=100, rows, ncu captures.
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2f_tests.log 2>&1; tail -2 $O/r2f_tests.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2f_bench_1gpu.json 2> $O/r2f_bench_1gpu.err
python bench.py --gpus 1 --steps 100 --warmup 5 --no-configs --no-cpu-baseline > $O/r2f_bench_1gpu_k100.json 2> /dev/null
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $O/r2f_reference_arm.json 2> /dev/null
python tools/bench_rows.py 20 > $O/r2f_rows.jsonl 2> /dev/null
bash tools/profile_r2.sh > $O/r2f_profile.log 2>&1
ls -la $O/r2f_* | head -20
