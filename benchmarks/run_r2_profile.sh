#!/bin/bash
# Round-2 evidence run on ONE B200: GPU tests, kernel benches, step bench (C2 / C3), ncu launch list and --set full captures
# of the kernels VERDICT r1 named.  Summaries (csv / logs) land in gpurun_out/; profiles/summarize.py turns them into profiles/*.md.
mkdir -p gpurun_out
M='dram__bytes_(read|write)\.sum|dram__cycles_active|gpu__dram_throughput|dram__throughput|sm__pipe_tensor_cycles_active|sm__warps_active|launch__registers_per_thread|gpu__time_duration.sum|sm__throughput|launch__grid_size|launch__block_size|sm__inst_executed_pipe_tensor|l1tex__data_pipe|smsp__inst_executed.sum|sm__pipe_fma_cycles_active|lts__t_bytes.sum'
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2_gpu_tests_1gpu.log; tail -4 gpurun_out/r2_gpu_tests_1gpu.log
timeout 300 python benchmarks/bench_kernels.py --what dot,multihot,gather128 > gpurun_out/r2_bench_kernels.log 2>&1; cut -c1-260 gpurun_out/r2_bench_kernels.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu_c2.log 2>&1; tail -1 gpurun_out/r2_bench_1gpu_c2.log | cut -c1-400
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu_c3.log 2>&1; tail -1 gpurun_out/r2_bench_1gpu_c3.log | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
cap() {  # name regex skip count cmd...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/r2_$name "$@" > gpurun_out/r2_ncu_$name.log 2>&1
  ncu -i gpurun_out/r2_$name.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,re
rd=list(csv.reader(sys.stdin))
if len(rd)<3: sys.exit(0)
hdr=rd[0]; pat=re.compile(r'$M')
for row in rd[2:]:
    print('## launch', row[hdr.index('ID')] if 'ID' in hdr else '?', row[hdr.index('Kernel Name')][:120] if 'Kernel Name' in hdr else '')
    for h,v in zip(hdr,row):
        if pat.search(h): print('  ',h,'=',v, rd[1][hdr.index(h)])
" > gpurun_out/r2_ncu_$name.txt
  head -c 1500 gpurun_out/r2_ncu_$name.txt
}
cap gemm_tc_cross_fwd gemm_tc_kernel 0 3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline
cap scatter_fast scatter_fast_kernel 1 1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline
cap dot 'dot_(fwd|bwd)_mma_kernel' 4 2 python benchmarks/bench_kernels.py --what dot --reps 2
cap multihot '(gather|scatter)_generic_kernel' 2 2 python benchmarks/bench_kernels.py --what multihot --reps 2
rm -f gpurun_out/r2_multihot.ncu-rep gpurun_out/r2_scatter_fast.ncu-rep   # keep the merge under 64 MiB: text summaries stay
ls -la gpurun_out | tail -20
