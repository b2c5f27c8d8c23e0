#!/usr/bin/env python
"""tools/sass_report.py [LIB]: evidence that the parity-critical arithmetic is never contracted.  Disassembles the library
(nvdisasm -g: SASS with source-line markers) and prints
  * the opcode histogram of the A1 coverage loop of tile_kernel<0,0,0> (FADD / FMUL only, no FFMA),
  * every source line that owns an FFMA anywhere in the library (all of them division / reciprocal expansions: the
    IEEE division of __fdiv_rn is a MUFU.RCP refined with FFMAs, which is how nvcc implements correctly rounded
    division; no FFMA may come from a fused a*b+c of the raster arithmetic)."""
import collections, os, re, subprocess, sys, tempfile
lib = sys.argv[1] if len(sys.argv) > 1 else "rusterizer_b200/librz_b200.so"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
func, cur, rows, sub = None, None, [], None
for ln in sass:
    m = re.match(r"\.text\.(\S+):", ln)
    if m: func = m.group(1); sub = None; continue
    m = re.match(r"\s*(\$\S+):", ln)   # local labels: the out-of-line slow path of the IEEE division is one of them
    if m and "slowpath" in m.group(1): sub = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m and func:
        ins = m.group(1).strip()
        op = [t for t in ins.split() if not t.startswith("@")][0].split(".")[0]
        rows.append((func, ("<" + re.sub(r"^\$__internal_\d+_\$", "", sub) + ">", 0) if sub else cur, op, ins))
src = open("rusterizer_b200/csrc/rz_tile.cuh").read().splitlines()
a1_lo = next(i + 1 for i, l in enumerate(src) if "phase A1: thread = (item" in l)
a1_hi = next(i + 1 for i, l in enumerate(src) if "S.nfrag keeps counting past the pool" in l)
print(f"== A1 coverage loop of tile_kernel<false,false,false> (rz_tile.cuh:{a1_lo}-{a1_hi}), SASS opcode histogram")
hist, last_tile, n = collections.Counter(), None, 0
for f, c, op, ins in rows:
    if "tile_kernelILb0ELb0ELb0" not in f: continue
    if c and c[0] == "rz_tile.cuh": last_tile = c[1]
    if last_tile and a1_lo <= last_tile < a1_hi:
        hist[op] += 1; n += 1
print(f"{n} instructions:", ", ".join(f"{k} {v}" for k, v in hist.most_common()))
print("FFMA in the A1 loop:", hist.get("FFMA", 0), "   (FADD", hist.get("FADD", 0), "FMUL", hist.get("FMUL", 0), ": one SASS op per source-level f32 operation)")
print()
print("== every source line that owns an FFMA, whole library (kernel instantiations merged)")
ff = collections.Counter()
for f, c, op, ins in rows:
    if op == "FFMA": ff[c] += 1
tot = sum(ff.values())
lines_cache = {}
def text(c):
    if c is None: return ""
    for root in ("rusterizer_b200/csrc/", "/usr/local/cuda/include/", "/usr/local/cuda/include/crt/"):
        p = root + c[0]
        if os.path.exists(p):
            L = lines_cache.setdefault(p, open(p, errors="ignore").read().splitlines())
            return L[c[1] - 1].strip()[:110] if c[1] <= len(L) else ""
    return ""
for c, v in ff.most_common():
    print(f"{v:6d}  {c[0] if c else '?'}:{c[1] if c else 0:<5d} {text(c)}")
print(f"total FFMA {tot}; lines outside fdiv()/division helpers:",
      sum(v for c, v in ff.items() if c and not ("fdiv" in text(c) or "slowpath" in c[0])))
