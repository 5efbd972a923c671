#!/bin/bash
# One gpurun call: GPU tests, the bench lines, ncu launch list and per-kernel captures.  usage: gpu_session.sh <tag> [parts...]
# (scratch outputs go to gpurun_out/<tag>_*; profiles/summarize.py turns them into the tracked summaries)
tag=${1:-c}; shift
parts=${*:-tests bench kln batch ncu}
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
O=gpurun_out/$tag
for p in $parts; do
  case $p in
    tests) timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > ${O}_tests.txt ;;
    bench) timeout 600 python bench.py > ${O}_bench.json 2> ${O}_bench.err ;;
    ref)   timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err ;;
    kln)   timeout 600 python bench.py --workload kln > ${O}_bench_kln.json 2> ${O}_bench_kln.err ;;
    batch) for b in 4096 8192; do timeout 300 python bench.py --batch $b --no-cpu-baseline > ${O}_bench_b$b.json 2> ${O}_bench_b$b.err; done ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > ${O}_launches.log 2>&1 ;;
    ncu)   timeout 900 ncu --set full --clock-control none --import-source on -k regex:"deposit_kernel|sample_collide|moments_kernel" -s 3 -c 3 -f -o ${O}_prof python profiles/prof_run.py 2048 2048 > ${O}_prof.log 2>&1 ;;
    ncukln) timeout 900 ncu --set full --clock-control none --import-source on -k regex:"deposit_kernel|combine_kernel|moments_kernel" -s 3 -c 3 -f -o ${O}_prof_kln python profiles/prof_run.py 2048 2048 kln > ${O}_prof_kln.log 2>&1 ;;
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.txt 2>&1 ;;
    scanexe) timeout 400 python bench.py --workload scan-exe --steps 2 --warmup 1 > ${O}_bench_scanexe.json 2> ${O}_bench_scanexe.err ;;
    ebe)   timeout 400 python bench.py --workload ebe --steps 2 --warmup 1 > ${O}_bench_ebe.json 2> ${O}_bench_ebe.err ;;
    avg)   timeout 600 python bench.py --workload avg --steps 2 --warmup 1 > ${O}_bench_avg.json 2> ${O}_bench_avg.err ;;
    others) for w in ppb auau sqrt nbd; do timeout 300 python bench.py --workload $w --cpu-sample-events 100 > ${O}_bench_$w.json 2> ${O}_bench_$w.err; done ;;
    sanit) for w in glb kln nbd sqrt; do timeout 600 compute-sanitizer --tool memcheck python profiles/sanitizer_run.py $w 2>&1 | tail -1 > ${O}_mem_$w.txt; timeout 700 compute-sanitizer --tool racecheck python profiles/sanitizer_run.py $w 2>&1 | tail -1 > ${O}_race_$w.txt; done ;;
    sass)  cuobjdump -sass supermc_b200/libsupermc_b200.so > ${O}_sass.txt 2>&1 ;;
  esac
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${O}_smi.txt 2>&1
ls -la gpurun_out | tail -30
