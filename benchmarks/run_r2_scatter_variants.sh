#!/bin/bash
# Multi-hot backward variants (lookups per pass HU, CTAs per SM MINB, grid cap in CTAs per SM) built as separate libraries
# (KRS_EXTRA_FLAGS="-DKRS_SCAT_HU=.. -DKRS_SCAT_MINB=.. -DKRS_SCAT_GRIDMUL=..") and selected with KRS_B200_LIB.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/r2_last_gpu_tests.log; tail -1 gpurun_out/r2_last_gpu_tests.log
{
echo "default (HU=4 MINB=4 grid cap 64/SM):"; timeout 60 python benchmarks/bench_kernels.py --what multihot 2>&1 | grep scatter | cut -c1-260
for f in variants_tmp/*.so; do echo "$f:"; KRS_B200_LIB=$PWD/$f timeout 60 python benchmarks/bench_kernels.py --what multihot 2>&1 | grep scatter | cut -c1-260; done
} > gpurun_out/r2_multihot_scatter_variants.txt 2>&1
cat gpurun_out/r2_multihot_scatter_variants.txt
M='dram__bytes_(read|write)\.sum$|gpu__dram_throughput.avg.pct|sm__warps_active.avg.pct|launch__registers_per_thread|gpu__time_duration.sum|launch__grid_size|lts__t_bytes.sum$'
timeout 200 ncu --set full --clock-control none -k regex:scatter_sample_kernel -s 2 -c 1 -f -o gpurun_out/ss python benchmarks/bench_kernels.py --what multihot --reps 2 > /dev/null 2>&1
ncu -i gpurun_out/ss.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,re
rd=list(csv.reader(sys.stdin))
if len(rd)<3: sys.exit(0)
hdr=rd[0]; pat=re.compile(r'$M')
for row in rd[2:]:
    print('## launch', row[hdr.index('Kernel Name')][:110])
    for h,v in zip(hdr,row):
        if pat.search(h): print('  ',h,'=',v, rd[1][hdr.index(h)])
" > gpurun_out/r2_ncu_scatter_sample.txt
rm -f gpurun_out/ss.ncu-rep; cat gpurun_out/r2_ncu_scatter_sample.txt
timeout 200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_frontend.py -q -p no:cacheprovider -k "scatter or multihot or embed or ragged or bwd" > gpurun_out/r2_sanitizer_memcheck_scatter_sample.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer_memcheck_scatter_sample.log | tail -3
