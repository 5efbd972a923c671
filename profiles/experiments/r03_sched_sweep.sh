#!/bin/bash
# scheduling sweep: occupancy caps (shared-memory padding) on sampler / moments, deposit CTAs per SM, pipeline slots
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}" || exit 1
mkdir -p gpurun_out
run() { # tag env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/s1_$tag.json 2> gpurun_out/s1_$tag.err
}
run base
run sp5 SMC_SAMPLE_PAD=2048
run sp4 SMC_SAMPLE_PAD=9216
run mp2 SMC_MOM_PAD=49152
run sp4mp2 SMC_SAMPLE_PAD=9216 SMC_MOM_PAD=49152
run sp4mp1 SMC_SAMPLE_PAD=9216 SMC_MOM_PAD=90000
run sp3mp2 SMC_SAMPLE_PAD=21000 SMC_MOM_PAD=49152
run d2sp4 SMC_DEP_CTAS=2 SMC_SAMPLE_PAD=9216
run d2mp1 SMC_DEP_CTAS=2 SMC_MOM_PAD=90000
run sl3 SMC_SLOTS=3
run sl2 SMC_SLOTS=2
run b1024 SMC_SLOTS=4 BATCH=1024
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s1_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms_per_step'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
