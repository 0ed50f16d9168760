#!/bin/bash
# ncu --set full of one cross-forward GEMM per engine (tests/tc_trace.py launches it 4 times; capture the 3rd)
mkdir -p gpurun_out
for e in tcgen05 tcgen05_ts; do
  ENGINE=$e KRS_TC_FUSE_N=${KRS_TC_FUSE_N:-1} timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 \
     -o gpurun_out/gemm_${e}_r1g -f python tests/tc_trace.py > gpurun_out/ncu_${e}.log 2>&1
  echo "ncu $e rc=$?"
done
ls -la gpurun_out/*.ncu-rep
