#!/bin/bash
# Round-2 final evidence run on ONE B200 (after the last kernel change): GPU tests, kernel benches, the driver-contract bench
# lines (C2 with the CPU baseline, C3), ncu launch list, --set full metrics of the kernels changed since run_r2_profile.sh,
# compute-sanitizer memcheck over the small-shape kernel tests.  Only text summaries are kept (the .ncu-rep files are dropped).
mkdir -p gpurun_out
M='dram__bytes_(read|write)\.sum$|dram__bytes_(read|write)\.sum\.per_second|gpu__dram_throughput.avg.pct|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__warps_active.avg.pct|launch__registers_per_thread|gpu__time_duration.sum|sm__throughput.avg.pct|launch__grid_size|launch__block_size|lts__t_bytes.sum$'
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/r2_final_gpu_tests.log; tail -3 gpurun_out/r2_final_gpu_tests.log | cut -c1-200
timeout 400 python benchmarks/bench_kernels.py --what dot,multihot,gather128,topk > gpurun_out/r2_final_bench_kernels.jsonl 2>&1; cut -c1-220 gpurun_out/r2_final_bench_kernels.jsonl
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_1gpu_c2.log 2>&1; tail -1 gpurun_out/r2_final_bench_1gpu_c2.log | cut -c1-300
timeout 200 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_final_bench_1gpu_c3.log 2>&1; tail -1 gpurun_out/r2_final_bench_1gpu_c3.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_final_bench_under_ncu.log 2>&1
cap() {  # name regex skip count cmd...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 300 ncu --set full --clock-control none -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/r2f_$name "$@" > gpurun_out/r2f_ncu_$name.log 2>&1
  ncu -i gpurun_out/r2f_$name.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,re
rd=list(csv.reader(sys.stdin))
if len(rd)<3: sys.exit(0)
hdr=rd[0]; pat=re.compile(r'$M')
for row in rd[2:]:
    print('## launch', row[hdr.index('Kernel Name')][:110] if 'Kernel Name' in hdr else '')
    for h,v in zip(hdr,row):
        if pat.search(h): print('  ',h,'=',v, rd[1][hdr.index(h)])
" > gpurun_out/r2_final_ncu_$name.txt
  rm -f gpurun_out/r2f_$name.ncu-rep gpurun_out/r2f_ncu_$name.log
  head -c 900 gpurun_out/r2_final_ncu_$name.txt
}
cap gather_scatter '(gather_fast|scatter_fast)_kernel' 2 2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline
cap multihot '(gather_sample|scatter_sample)_kernel' 2 2 python benchmarks/bench_kernels.py --what multihot --reps 2
cap dot 'dot_(fwd|bwd)_mma_kernel' 6 2 python benchmarks/bench_kernels.py --what dot --reps 2
timeout 280 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_exchange.py tests/test_retrieval_helpers.py -q -p no:cacheprovider -x -k "gather or scatter or route or slot or compact or row_topk or hard_negative or accidental or multihot or embed" > gpurun_out/r2_final_sanitizer_memcheck_kernels.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_final_sanitizer_memcheck_kernels.log | tail -3
ls -la gpurun_out | grep r2_final
