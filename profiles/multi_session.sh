#!/bin/bash
# multi-GPU evidence on N GPUs of one box: usage multi_session.sh <tag> <N>   (outputs gpurun_out/<tag>_*)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
tag=$1; N=$2; mkdir -p gpurun_out; O=gpurun_out/$tag
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29701 bench.py --gpus $N > ${O}_bench_n$N.json 2> ${O}_bench_n$N.err
timeout 300 $TR --master-port 29702 bench.py --gpus $N --workload kln > ${O}_bench_kln_n$N.json 2> ${O}_bench_kln_n$N.err
timeout 300 $TR --master-port 29703 bench.py --gpus $N --workload avg --steps 4 --warmup 1 > ${O}_bench_avg_n$N.json 2> ${O}_bench_avg_n$N.err
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 > ${O}_multi_gpu_tests_n$N.txt
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > ${O}_smi_n$N.txt
python - <<PY
import json
for w in ["bench","bench_kln","bench_avg"]:
    try:
        d=json.loads(open("${O}_%s_n$N.json"%w).read().strip().splitlines()[-1]); print(w, d["n_gpus"], round(d["value"]), round(d["ms_per_step"],2), d["config"].get("allreduce_ms"), d["config"].get("collective"))
    except Exception as e: print(w,"ERR",e)
PY
cat ${O}_multi_gpu_tests_n$N.txt
