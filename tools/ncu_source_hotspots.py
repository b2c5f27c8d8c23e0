#!/usr/bin/env python
"""tools/ncu_source_hotspots.py REPORT.ncu-rep [KERNEL_REGEX] [TOP]: warp instructions executed and stall samples per
CUDA source line of one kernel of an `ncu --set full --import-source on` report (built with -lineinfo)."""
import csv, subprocess, sys
rep = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else "tile_kernel"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, agg, first_kernel = None, {}, None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        if first_kernel is None: first_kernel = r[1]
        elif r[1] != first_kernel: break   # only the first matching launch
        continue
    if r[0] == 'Kernel Name': continue
    if r[0] and r[0].isdigit():
        try: inst, samp = int(r[7]), int(r[4])
        except Exception: continue
        k = (cur, int(r[0]))
        a = agg.get(k, (0, 0, r[1][:120]))
        agg[k] = (a[0] + inst, a[1] + samp, a[2])
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(first_kernel); print('total inst (inlined frames counted per line)', tot, 'samples', tots)
byfile = {}
for (f, l), v in agg.items(): byfile[f] = byfile.get(f, 0) + v[0]
print({k: v for k, v in sorted(byfile.items(), key=lambda kv: -kv[1])[:6]})
for (f, l), (i, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{l:4d} inst {i:9d} {100*i/tot:5.1f}% samp {100*s/max(tots,1):5.1f}%  {src}")
