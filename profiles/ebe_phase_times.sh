#!/bin/bash
# phase times of operation 1 through the executable (SMC_TIMING=1): text, raw-binary and sd+ed output, each twice
ROOT="${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}"
mkdir -p /tmp/sx; cp "$ROOT/supermc_b200/parameters.dat" /tmp/sx/; cd /tmp/sx
for rep in 1 2; do for extra in "use_ed=0" "use_ed=0 output_binary=1" "use_ed=1"; do
  rm -rf data; mkdir data; sync; echo "== $extra (run $rep)"
  ( time SMC_TIMING=1 "$ROOT/supermc_b200/superMC_b200.e" which_mc_model=5 sub_model=1 Aproj=197 Atarg=197 ecm=200 alpha=0.14 cc_fluctuation_model=6 cc_fluctuation_Gamma_theta=0.61 maxx=13 maxy=13 finalFactor=1 operation=1 nev=1000 use_sd=1 use_block=1 use_4col=0 randomSeed=9 $extra > /dev/null ) 2>&1 | grep -v "QuarkPos\|smc_create\|^$\|user\|sys"
done; done
