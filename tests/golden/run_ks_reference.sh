#!/bin/bash
# Statistical reference samples for the KS tests (build container only): 8 processes of the UNMODIFIED reference binary
# (oracle/_ref/superMC_ref.e, its own 8-process mode, seeds 1000..1007, 12,500 accepted events each, operation 9, 261^2)
# per system, under /tmp/ks_<system>_<i>/.  tests/golden/make_ks_reference.py <system> then stores the columns.
# usage: run_ks_reference.sh <system> ...      systems: pbpb2760 auau200 ppb5020 auau200_kln
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
REF=$ROOT/oracle/_ref
COMMON="maxx=13 maxy=13 dx=0.1 dy=0.1 operation=9 finalFactor=1 bmin=0 bmax=20 Npmin=2 Npmax=500 shape_of_nucleons=2 collision_criterion=2 shape_of_entropy=2 ecc_from_order=1 ecc_to_order=9 use_sd=1 use_ed=0 nev=12500"
for sys in "$@"; do
  case $sys in
    pbpb2760)    P="which_mc_model=5 sub_model=1 Aproj=208 Atarg=208 ecm=2760 alpha=0.118 cc_fluctuation_model=6 cc_fluctuation_Gamma_theta=0.75" ;;
    auau200)     P="which_mc_model=5 sub_model=1 Aproj=197 Atarg=197 ecm=200 alpha=0.14 cc_fluctuation_model=6 cc_fluctuation_Gamma_theta=0.61" ;;
    ppb5020)     P="which_mc_model=5 sub_model=1 Aproj=1 Atarg=208 ecm=5020 alpha=0.118 cc_fluctuation_model=6 cc_fluctuation_Gamma_theta=0.75" ;;
    auau200_kln) P="which_mc_model=1 sub_model=7 lambda=0.218 Aproj=197 Atarg=197 ecm=200 cc_fluctuation_model=0 tmax=71 tmax_subdivision=3" ;;
    *) echo "unknown system $sys"; exit 1 ;;
  esac
  for i in 0 1 2 3 4 5 6 7; do
    d=/tmp/ks_${sys}_$i; rm -rf $d; mkdir -p $d/data
    for f in parameters.dat EOS tables; do ln -s $REF/run_zero/$f $d/$f; done
    (cd $d && $REF/superMC_ref.e $COMMON $P randomSeed=$((1000 + i)) > log.txt 2>&1) &
  done
  wait
  echo "$sys done: $(cat /tmp/ks_${sys}_*/data/sn_ecc_eccp_10.dat | wc -l) rows"
done
