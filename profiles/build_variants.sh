#!/bin/bash
# builds deposit-kernel tuning variants next to the product library (profiling aid, not part of the product)
cd "$(dirname "$0")/../supermc_b200/csrc" || exit 1
i=0
for v in "$@"; do
  i=$((i+1))
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $v -dc -c smc_grid.cu -o /tmp/smc_grid_v$i.o || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libsupermc_b200_v$i.so smc_api.o smc_sample.o /tmp/smc_grid_v$i.o smc_kln.o smc_avg.o smc_comm.o smc_profile3d.o -Xcompiler -fPIC -ldl || exit 1
  echo "v$i: $v"
done
