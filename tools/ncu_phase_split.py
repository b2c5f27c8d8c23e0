#!/usr/bin/env python
"""tools/ncu_phase_split.py REPORT.ncu-rep: warp instructions of the tile kernel per phase.  SASS instructions that ncu
attributes to inlined helpers (rz_exact.cuh ...) are assigned to the phase of the nearest preceding instruction (in
address order) that maps to rz_tile.cuh; phases are line ranges found from the marker comments in the source."""
import csv, re, subprocess, sys
rep = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else "tile_kernel"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
src = open("rusterizer_b200/csrc/rz_tile.cuh").read().splitlines()
def find(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]: return i + 1
    raise KeyError(pat)
k0 = find(") tile_kernel(FrameParams P) {")
marks = [("prologue", k0), ("tile top / steal", find("for (;;) {", k0)), ("sort + entry load (A0 head)", find("if (n > 0 && !fast_done) {", k0)),
         ("wild walk", find("run of items for the literal walk", k0)), ("A0 table + scan", find("chunk of small items", k0)),
         ("direct/budget", find("chunk dominated by large items: pixel-parallel, deferred", k0)),
         ("A1 coverage+depth", find("phase A1: thread = (item", k0)), ("A1 tail/retry", find("S.nfrag keeps counting past the pool", k0)), ("A2 depths", find("phase A2: thread = fragment", k0)),
         ("B replay", find("phase B: thread = pixel", k0)), ("C shade", find("phase C: thread = fragment", k0)),
         ("resolve+store+clear", find("every path through the chunk loop ends with a barrier", k0)), ("epilogue", find("the rest of the tiles nothing was binned into", k0))]
helpers = [("shade()", find("__device__ __forceinline__ uint32_t shade("), k0)]
def phase_of(line):
    if line < k0:
        return "fn:" + ("sort" if find("block_sort(T *a")-2 <= line < find("struct __align__(16) BigSetup") or find("sort_tile_list(SM &S") <= line < find("Stage the records of") else
                        "shade/texture" if line < find("Bitonic network") else "direct_chunk/stage" if line < find("Tiles nothing was binned into") else "clear_empty" if line < find("Tiles whose whole list has at most FAST_N items") else "fast_tile")
    cur = marks[0][0]
    for name, l in marks:
        if line >= l: cur = name
    return cur
cur_file, seen_fn, inst = None, None, []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        if seen_fn is None: seen_fn = r[1]
        elif r[1] != seen_fn: break
        continue
    if r[0].isdigit(): cur_line = int(r[0]); continue
    if r[0] == '' and len(r) > 8 and r[2].startswith('0x'):
        try: inst.append((int(r[2], 16), cur_file, cur_line, int(r[7]), int(r[8]), int(r[4] or 0)))
        except ValueError: pass
# an address may be listed under several source lines (inlining): keep the rz_tile.cuh attribution when there is one
best = {}
for a, f, l, wi, ti, sm in inst:
    if a not in best or (f == 'rz_tile.cuh' and best[a][0] != 'rz_tile.cuh'): best[a] = (f, l, wi, ti, sm)
acc, cur = {}, "prologue"
for a in sorted(best):
    f, l, wi, ti, sm = best[a]
    if f == 'rz_tile.cuh': cur = phase_of(l)
    w, t, s = acc.get(cur, (0, 0, 0)); acc[cur] = (w + wi, t + ti, s + sm)
tot = sum(v[0] for v in acc.values()); tots = sum(v[2] for v in acc.values())
print(seen_fn, "warp inst", tot)
for k, (w, t, s) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:32s} warp inst {w:9d} {100*w/tot:5.1f}%  lanes/inst {t/max(w,1):5.1f}  stall samples {100*s/max(tots,1):5.1f}%")
