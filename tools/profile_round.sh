#!/bin/bash
# tools/profile_round.sh TAG: the round's profiling passes on one B200 (run under gpurun).  Writes to gpurun_out/:
#   launches_TAG.csv      ncu launch list (gpu__time_duration.sum, --clock-control none) of a short bench run
#   prof_TAG.ncu-rep      one `--set full` capture of every kernel of one C2 frame (source imported)
#   prof_TAG_c3.ncu-rep   the same for the tile kernel of the C3 frame (250K near-clipped triangles at 4K)
#   bench_TAG.json/.err   the unprofiled bench line (with cpu_baseline), ref_TAG.json the --impl reference line
#   tile_times_TAG.txt, configs_TAG.jsonl   per-tile timing of the C2 frame, every BASELINE config on one GPU
# The captures come first: bench.py quotes profiles/traffic.json (ncu bytes and warp instructions per kernel), which
# tools/make_profiles.py TAG --traffic-only regenerates on the box from the fresh capture; run tools/make_profiles.py
# TAG afterwards (CPU side) to refresh everything under profiles/.
tag=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --inflight 1 --min-timed-s 0 > gpurun_out/ncu_launch_bench_$tag.log 2>&1
# full capture: skip the warm-up frames (6 kernels per frame: vertex, geom, clip, large_bin, order, tile)
ncu --set full --clock-control none --import-source on -k regex:'vertex_kernel|geom_kernel|clip_kernel|large_bin_kernel|order_kernel|tile_kernel' \
    --launch-skip 30 --launch-count 6 -f -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 3 --no-cpu-baseline --inflight 1 --min-timed-s 0 \
    > gpurun_out/ncu_full_$tag.log 2>&1
RZ_SCENE=c3 ncu --set full --clock-control none --import-source on -k regex:tile_kernel --launch-skip 8 --launch-count 1 -f \
    -o gpurun_out/prof_${tag}_c3 python tools/stage_times.py --child > gpurun_out/ncu_full_${tag}_c3.log 2>&1
python tools/make_profiles.py $tag --traffic-only > gpurun_out/traffic_$tag.log 2>&1
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/ref_$tag.json 2>> gpurun_out/bench_$tag.err
python tools/tile_times.py > gpurun_out/tile_times_$tag.txt 2>&1
python tools/run_configs.py > gpurun_out/configs_$tag.jsonl 2> gpurun_out/configs_$tag.err
ls -la gpurun_out | tail -12
tail -c 600 gpurun_out/bench_$tag.json
