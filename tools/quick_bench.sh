#!/bin/bash
# tools/quick_bench.sh TAG [LIB ...]: C2 bench (no CPU baseline) for the default library and each variant library;
# prints value / stage times, keeps the JSON lines in gpurun_out/quick_TAG.jsonl
tag=$1; shift
mkdir -p gpurun_out
for lib in default "$@"; do
  if [ "$lib" = default ]; then unset RZ_B200_LIB; else export RZ_B200_LIB=$PWD/$lib; fi
  python bench.py --no-cpu-baseline --steps 40 --warmup 5 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); k = j['roofline']['kernel_ms_all']
        print('$lib', 'ms/frame %.4f' % j['ms_per_step'], 'Mtris/s %.0f' % j['value'], {a: round(b, 4) for a, b in k.items()}, 'e2e ms %.3f' % j['e2e']['ms_per_step'])
        j['lib'] = '$lib'; open('gpurun_out/quick_$tag.jsonl', 'a').write(json.dumps(j) + '\n')
"
done
