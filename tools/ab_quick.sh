#!/bin/bash
# tools/ab_quick.sh [LIB ...]: GPU parity tests of the in-tree library, then per-stage times of the in-tree library and
# of each variant (build/*.so) on the scenes in AB_SCENES, then the C2 throughput line of each.
libs="$@"
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 )
for sc in ${AB_SCENES:-c2 overdraw c3 c4i c4ii c1}; do RZ_SCENE=$sc timeout 600 python tools/stage_times.py $libs; done
for lib in default $libs; do
  if [ "$lib" = default ]; then unset RZ_B200_LIB; else export RZ_B200_LIB=$PWD/$lib; fi
  python bench.py --no-cpu-baseline --steps 60 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('$lib', 'ms/frame %.4f Mtris/s %.0f latency %.4f e2e %.3f' % (j['ms_per_step'], j['value'], j.get('frame_latency_ms') or 0, j['e2e']['ms_per_step']))
"
done
