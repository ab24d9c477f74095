cd $GRAFT_REPO_ROOT
run() { name=$1; cpus=$2; shift 2; AVRF_BENCH_TRACE=1 taskset -c 0-$cpus python bench.py --no-configs --no-cpu-baseline "$@" > gpurun_out/r3_T_$name.json 2> gpurun_out/r3_T_$name.err; python - <<P
import json
for l in open("gpurun_out/r3_T_$name.json"):
    if l.startswith("{"):
        d=json.loads(l); c=d["config"]; print("$name", round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), c["concurrency"], c["own_thread_hashes"], c["mb_sha512_threads"])
P
}
run c16 15
run c12 11
run c8 7
run c4 3
run c16k100 15 --steps 100
run c16k50 15 --steps 50
