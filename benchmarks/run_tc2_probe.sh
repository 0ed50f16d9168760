#!/bin/bash
# GPU experiment: validate + time + trace the tcgen05 GEMM engines (SS: tcgen05, A-in-TMEM: tcgen05_ts), fused-N on/off.
mkdir -p gpurun_out
for f in 0 1; do
  KRS_TC_FUSE_N=$f KRS_TEST_TC_ENGINES=tcgen05,tcgen05_ts timeout 420 python -m pytest tests/test_gpu_tc.py -q --timeout 150 -p no:cacheprovider > gpurun_out/tc_tests_fuse$f.log 2>&1
  echo "tests fuse=$f rc=$?"; tail -2 gpurun_out/tc_tests_fuse$f.log
done
: > gpurun_out/gemm_probe.log
for f in 0 1; do KRS_TC_FUSE_N=$f timeout 200 python benchmarks/gemm_probe.py --engines tcgen05,tcgen05_ts >> gpurun_out/gemm_probe.log 2>&1; done
for e in tcgen05 tcgen05_ts; do for f in 0 1; do
  ENGINE=$e KRS_TC_FUSE_N=$f timeout 120 python tests/tc_trace.py > gpurun_out/trace_${e}_fuse$f.txt 2>&1
done; done
ENGINE=tcgen05 MODE=sgemm timeout 120 python tests/tc_trace.py > gpurun_out/trace_tcgen05_sgemm.txt 2>&1
cat gpurun_out/gemm_probe.log | cut -c1-200
