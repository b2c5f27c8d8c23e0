#!/bin/bash
# tools/ab_round.sh TAG [LIB ...]: one GPU call that (1) runs the GPU parity tests on the in-tree library,
# (2) compares per-stage times of the in-tree library with the variant libraries (build/*.so from
# tools/build_variant.sh / build_rev.sh) on the BASELINE scenes, (3) runs the C2 throughput bench for each and
# (4) sweeps the frames in flight (AB_SCENES / AB_INFLIGHT select scenes and sweep points).  Everything lands in
# gpurun_out/ab_TAG.*
tag=$1; shift
libs="$@"
mkdir -p gpurun_out
out=gpurun_out/ab_$tag.txt
: > $out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/ab_${tag}_tests.log
echo "== tests ==" >> $out; tail -3 gpurun_out/ab_${tag}_tests.log >> $out
for sc in ${AB_SCENES:-c2 c4ii c3 c1 overdraw c4i}; do
  echo "== stage times $sc ==" >> $out
  RZ_SCENE=$sc timeout 600 python tools/stage_times.py $libs >> $out 2>&1
done
echo "== C2 throughput (3 frames in flight) ==" >> $out
timeout 900 tools/quick_bench.sh $tag $libs >> $out 2>&1
echo "== frames in flight sweep (in-tree library) ==" >> $out
for k in ${AB_INFLIGHT:-1 2 4 6}; do
  timeout 300 python bench.py --no-cpu-baseline --steps 60 --warmup 5 --inflight $k 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('inflight $k ms/frame %.4f Mtris/s %.0f' % (j['ms_per_step'], j['value']))
" >> $out
done
cat $out
