cd $GRAFT_REPO_ROOT
run() { name=$1; shift; AVRF_BENCH_TRACE=1 taskset -c 0-3 python bench.py --no-configs --no-cpu-baseline "$@" > gpurun_out/r3_p4_$name.json 2> gpurun_out/r3_p4_$name.err; python - <<P
import json
for l in open("gpurun_out/r3_p4_$name.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$name", d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["concurrency"], d["config"]["mb_sha512_threads"])
P
}
run default
run h3c12 --hashers 3 --concurrency 12
run h3c24 --hashers 3 --concurrency 24
run h2c16 --hashers 2 --concurrency 16
run h0c4 --hashers 0 --concurrency 4
run default_k100 --steps 100
