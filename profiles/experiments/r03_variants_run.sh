#!/bin/bash
# bench every libsupermc_b200_v*.so built by profiles/build_variants.sh next to the product library; quick parity check first
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}" || exit 1
mkdir -p gpurun_out
tag=${1:-v}
for so in supermc_b200/libsupermc_b200.so supermc_b200/libsupermc_b200_v*.so; do
  n=$(basename $so .so | sed 's/libsupermc_b200//;s/^_//'); n=${n:-base}
  SMC_LIB=$PWD/$so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2 > gpurun_out/${tag}_${n}_parity.txt
  SMC_LIB=$PWD/$so timeout 200 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/${tag}_${n}.json 2> gpurun_out/${tag}_${n}.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${tag}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms_per_step'].items()}, open(f.replace('.json','_parity.txt')).read().strip().splitlines()[-1])
    except Exception as e: print(f, 'ERR', e)
PY
