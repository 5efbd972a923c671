#!/bin/bash
# BASELINE.json configs[0]: MC-Glauber Au+Au 200 GeV, 261x261, event-by-event entropy density + eccentricities (operation 1,
# use_sd=1 use_block=1), text output included -- this repo's executable vs the unmodified reference binary on the same box.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
R=$PWD; O=$R/gpurun_out/ebe
ARGS="which_mc_model=5 sub_model=1 Aproj=197 Atarg=197 ecm=200 alpha=0.14 cc_fluctuation_model=6 cc_fluctuation_Gamma_theta=0.61 maxx=13 maxy=13 dx=0.1 dy=0.1 finalFactor=1 operation=1 use_sd=1 use_ed=0 use_block=1 use_4col=0 randomSeed=9"
now() { date +%s.%N; }
d=$(mktemp -d); mkdir $d/data; cp supermc_b200/parameters.dat $d/
( cd $d; t0=$(now); $R/supermc_b200/superMC_b200.e $ARGS nev=8 > /dev/null; t1=$(now); rm -f data/*
  $R/supermc_b200/superMC_b200.e $ARGS nev=1000 > /dev/null; t2=$(now)
  echo "ours: start-up+8 events $(python3 -c "print(round($t1 - $t0, 2))") s; 1000 events $(python3 -c "print(round($t2 - $t1, 2))") s; files $(ls data | wc -l); MB $(du -sm data | cut -f1)" > $O.ours.txt )
rm -rf $d
if [ -x oracle/_ref/superMC_ref.e ]; then
  d=$(mktemp -d); mkdir $d/data; for f in parameters.dat EOS tables; do ln -s $R/oracle/_ref/run_zero/$f $d/$f; done
  ( cd $d; t0=$(now); $R/oracle/_ref/superMC_ref.e $ARGS nev=1 > /dev/null; t1=$(now); rm -f data/*
    $R/oracle/_ref/superMC_ref.e $ARGS nev=41 > /dev/null; t2=$(now)
    echo "reference: start-up+1 event $(python3 -c "print(round($t1 - $t0, 2))") s; 41 events $(python3 -c "print(round($t2 - $t1, 2))") s; files $(ls data | wc -l); MB $(du -sm data | cut -f1)" > $O.ref.txt )
  rm -rf $d
fi
cat $O.ours.txt $O.ref.txt
