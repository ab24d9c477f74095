#!/bin/bash
# usage: final_ngpu.sh N   - the driver's bench command on N GPUs (+ the GPU tests that need several devices)
cd $GRAFT_REPO_ROOT
N=$1
O=gpurun_out
python -m pytest tests/test_gpu_dist.py tests/test_gpu_cpp.py tests/test_gpu_round2.py -m gpu -x -q -k "dist or sharded or nccl or cpp" > $O/r2f_tests_${N}gpu.log 2>&1; tail -2 $O/r2f_tests_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2f_bench_${N}gpu.json 2> $O/r2f_bench_${N}gpu.err
tail -c 300 $O/r2f_bench_${N}gpu.err; wc -c $O/r2f_bench_${N}gpu.json; nproc
