#!/bin/bash
# Round-end evidence on one B200: full GPU test suite, smoke, bench (with the CPU baseline), ncu launch list of the
# step, ncu --set full of the dominant GEMM and of the tensor-pipe top-k kernel, per-kernel benches (C3 / C4).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/gpu_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_1gpu.log 2>&1; echo "bench rc=$?"
timeout 300 python bench.py --optimizer adagrad --no-cpu-baseline > gpurun_out/bench_1gpu_adagrad.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
ENGINE=tcgen05_ts timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 \
    -o gpurun_out/gemm_tc_ts_final -f python tests/tc_trace.py > gpurun_out/ncu_gemm_final.log 2>&1; echo "ncu gemm rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:topk_tc_kernel -s 1 -c 1 \
    -o gpurun_out/topk_tc_final -f python benchmarks/topk_probe.py --engines tcgen05 --nc 2000000 --reps 1 > gpurun_out/ncu_topk_final.log 2>&1; echo "ncu topk rc=$?"
timeout 300 python benchmarks/topk_probe.py --engines tcgen05,ffma --reps 2 > gpurun_out/topk_probe.log 2>&1
timeout 300 python benchmarks/bench_kernels.py --what dot,gather128 > gpurun_out/bench_kernels.log 2>&1
timeout 200 python benchmarks/gemm_probe.py --engines tcgen05,tcgen05_ts > gpurun_out/gemm_probe.log 2>&1
ENGINE=tcgen05_ts timeout 120 python tests/tc_trace.py > gpurun_out/trace_tcgen05_ts_final.txt 2>&1
tail -1 gpurun_out/bench_1gpu.log | cut -c1-300; cat gpurun_out/topk_probe.log
