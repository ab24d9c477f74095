cd $GRAFT_REPO_ROOT
run() { name=$1; cpus=$2; shift 2; taskset -c 0-$cpus python bench.py --no-configs --no-cpu-baseline "$@" > gpurun_out/r3_T_$name.json 2> gpurun_out/r3_T_$name.err; python - <<P
import json
for l in open("gpurun_out/r3_T_$name.json"):
    if l.startswith("{"):
        d=json.loads(l); c=d["config"]; print("$name", round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), c["concurrency"], c["own_thread_hashes"], c["mb_sha512_threads"])
P
}
run c16own 15 --hashers 0 --concurrency 20
run c16mix 15
run c12own 11 --hashers 0 --concurrency 20
run c12mix 11
run c8own 7 --hashers 0 --concurrency 20
