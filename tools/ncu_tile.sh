#!/bin/bash
# tools/ncu_tile.sh TAG [LIB]: one `ncu --set full` capture of a warmed-up C2 tile kernel (and geometry kernels) -> gpurun_out/prof_TAG.ncu-rep
tag=$1; lib=$2
mkdir -p gpurun_out
if [ -n "$lib" ]; then export RZ_B200_LIB=$PWD/$lib; fi
ncu --set full --clock-control none --import-source on -k regex:'vertex_kernel|geom_kernel|clip_kernel|large_bin_kernel|order_kernel|tile_kernel|mid_kernel' \
    --launch-skip ${NCU_SKIP:-30} --launch-count ${NCU_COUNT:-6} -f -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 3 --no-cpu-baseline --inflight 1 \
    > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
