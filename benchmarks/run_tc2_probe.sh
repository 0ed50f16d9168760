#!/bin/bash
# GPU experiment: validate + time + trace the tcgen05 GEMM engines (SS: tcgen05, A-in-TMEM: tcgen05_ts)
mkdir -p gpurun_out
export KRS_TC_FUSE_N=${KRS_TC_FUSE_N:-1}
for b in 0 1; do
  KRS_TC_B_LO_TMA=$b KRS_TEST_TC_ENGINES=tcgen05,tcgen05_ts timeout 420 python -m pytest tests/test_gpu_tc.py -q --timeout 150 -p no:cacheprovider > gpurun_out/tc_tests_blo$b.log 2>&1
  echo "tests b_lo_tma=$b rc=$?"; tail -2 gpurun_out/tc_tests_blo$b.log
done
: > gpurun_out/gemm_probe.log
for b in 0 1; do echo "# KRS_TC_B_LO_TMA=$b" >> gpurun_out/gemm_probe.log; KRS_TC_B_LO_TMA=$b timeout 200 python benchmarks/gemm_probe.py --engines tcgen05,tcgen05_ts >> gpurun_out/gemm_probe.log 2>&1; done
for e in tcgen05 tcgen05_ts; do
  ENGINE=$e timeout 120 python tests/tc_trace.py > gpurun_out/trace_${e}_blo.txt 2>&1
done
cat gpurun_out/gemm_probe.log | cut -c1-200
