#!/bin/bash
# Branch wip/round2 = main + wip/ksub2 (tcgen05_ts ring stages of 2 x 16 k) + wip/topk-clo (candidate lo plane by TMA)
# + trace compile-out + one-add lo rounding + rotated top-k windows + krs_adamw_cold (decay-only sweep of never-touched rows)
# + pipelined AdamW (krs_adamw_rows / krs_adamw_skip, DCN.train_on_batch_pipelined, bench.py --pipeline-adamw)
# + keras_rs_b200/dlrm.py (ml_perf / C3 model on the public layers, oracle in np_oracle.dlrm_forward/backward).
# Neither has run on a GPU yet.  Build first (python -c "import __graft_entry__ as g; g.build()"), then:
#   gpurun --timeout 1500 -- 'bash benchmarks/run_wip_validation.sh'
mkdir -p gpurun_out
KRS_TEST_TC_ENGINES=tcgen05,tcgen05_ts timeout 420 python -m pytest tests/test_gpu_tc.py -q --timeout 150 -p no:cacheprovider > gpurun_out/wip_tc_tests.log 2>&1
echo "tc tests rc=$?"; tail -3 gpurun_out/wip_tc_tests.log
timeout 420 python -m pytest tests/test_gpu_kernels.py -q -k "topk or brute or retrieval" --timeout 150 -p no:cacheprovider > gpurun_out/wip_topk_tests.log 2>&1
echo "topk tests rc=$?"; tail -3 gpurun_out/wip_topk_tests.log
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -k "adamw or graph or dcn" --timeout 150 -p no:cacheprovider > gpurun_out/wip_adamw_tests.log 2>&1
echo "adamw/model tests rc=$?"; tail -3 gpurun_out/wip_adamw_tests.log
timeout 300 python -m pytest tests/test_gpu_zz_dlrm.py -q --timeout 150 -p no:cacheprovider > gpurun_out/wip_dlrm_tests.log 2>&1
echo "dlrm tests rc=$?"; tail -3 gpurun_out/wip_dlrm_tests.log
timeout 200 python benchmarks/gemm_probe.py --engines tcgen05,tcgen05_ts > gpurun_out/wip_gemm_probe.log 2>&1; cut -c1-170 gpurun_out/wip_gemm_probe.log
# (tests/tc_trace.py needs a -DKRS_TC_TRACE=1 build on this branch; skipped here)
# tile order: rotated windows (default on this branch) vs plain order (KRS_TOPK_WIN=0)
timeout 300 python benchmarks/topk_probe.py --engines tcgen05 > gpurun_out/wip_topk_probe.log 2>&1; cat gpurun_out/wip_topk_probe.log
KRS_TOPK_WIN=0 timeout 300 python benchmarks/topk_probe.py --engines tcgen05 > gpurun_out/wip_topk_probe_win0.log 2>&1; cat gpurun_out/wip_topk_probe_win0.log
timeout 300 python - > gpurun_out/wip_topk_lo_probe.log 2>&1 <<'PY'
import json, os, sys, torch
sys.path.insert(0, os.getcwd())
import keras_rs_b200 as K
g = torch.Generator(device="cuda").manual_seed(42)
C = torch.randn((10_000_000, 64), device="cuda", generator=g); Q = torch.randn((4096, 64), device="cuda", generator=g)
lo = K.ops.split_candidates_lo(C)
K.ops.top_k_scores(Q, C, None, 100, cand_lo=lo); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): s, i = K.ops.top_k_scores(Q, C, None, 100, cand_lo=lo)
e1.record(); torch.cuda.synchronize()
s0, i0 = K.ops.top_k_scores(Q, C, None, 100)
print(json.dumps(dict(engine="tcgen05 + precomputed C_lo", ms=round(e0.elapsed_time(e1) / 3, 3), same_as_in_kernel_split=bool(torch.equal(s, s0) and torch.equal(i, i0)))))
PY
cat gpurun_out/wip_topk_lo_probe.log
timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/wip_bench.log 2>&1; tail -1 gpurun_out/wip_bench.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --pipeline-adamw > gpurun_out/wip_bench_pipelined.log 2>&1; tail -1 gpurun_out/wip_bench_pipelined.log | cut -c1-200
