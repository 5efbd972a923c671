cd "${GRAFT_REPO_ROOT}" || exit 1
N=8; O=gpurun_out/r02; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29701 bench.py --gpus $N > ${O}_bench_n$N.json 2> ${O}_bench_n$N.err
timeout 300 $TR --master-port 29703 bench.py --gpus $N --workload avg --steps 6 --warmup 1 --no-cpu-baseline > ${O}_bench_avg_n$N.json 2> ${O}_bench_avg_n$N.err
timeout 300 $TR --master-port 29702 bench.py --gpus $N --workload kln --no-cpu-baseline > ${O}_bench_kln_n$N.json 2> ${O}_bench_kln_n$N.err
python - <<PY
import json
for w in ["bench","bench_kln","bench_avg"]:
    try:
        d=json.loads(open("${O}_%s_n$N.json"%w).read().strip().splitlines()[-1]); print(w, d["n_gpus"], round(d["value"]), round(d["ms_per_step"],2), d["config"].get("allreduce_ms"), d["config"].get("step_ms"))
    except Exception as e: print(w,"ERR",e)
PY
