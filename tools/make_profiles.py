#!/usr/bin/env python
"""tools/make_profiles.py TAG: turn the outputs of tools/profile_round.sh (gpurun_out/*_TAG.*) into the tracked
summaries under profiles/ (r01_launches.csv + summary, r01_ncu_full_summary.txt, traffic.json, tile times, configs,
bench lines)."""
import collections, csv, json, shutil, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
RND = "r02"  # prefix of the tracked files under profiles/
TRAFFIC_ONLY = "--traffic-only" in sys.argv  # on the GPU box, between the ncu capture and the bench run: bench.py quotes
G, P = "gpurun_out/", "profiles/"             # profiles/traffic.json, which has to describe the code being measured
if not TRAFFIC_ONLY:
    rows = list(csv.reader(open(f"{G}launches_{tag}.csv", errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    H = rows[hi]; kn, mv, mu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    d = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv: continue
        v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1.0)
        d.setdefault(r[kn][:70], []).append(v)
    ours = collections.OrderedDict()  # template instantiations of one kernel (tile_kernel<DBG, EXT, DIRECT>) count as one
    for k, v in d.items():
        if "rz::" in k:
            ours.setdefault(k.split("<")[0].split("(")[0].replace("void ", ""), []).extend(v)
    tot = sum(sum(v) / len(v) for v in ours.values())
    out = [f"# ncu launch list summary ({RND})  --  source: profiles/{RND}_launches.csv",
           "# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --inflight 1 --min-timed-s 0",
           "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes", "",
           f"{'kernel':70s} {'launches':>8s} {'mean_us':>9s} {'min_us':>9s} {'max_us':>9s}"]
    out += [f"{k:70s} {len(v):8d} {sum(v)/len(v):9.2f} {min(v):9.2f} {max(v):9.2f}" for k, v in d.items()]
    out += ["", "share of one frame (our kernels, mean launch time):"]
    out += [f"  {k:62s} {sum(v)/len(v)/tot*100:5.1f}%" for k, v in ours.items()]
    open(P + f"{RND}_launches_summary.txt", "w").write("\n".join(out) + "\n")
    shutil.copy(f"{G}launches_{tag}.csv", P + f"{RND}_launches.csv")
open(P + f"{RND}_ncu_full_summary.txt", "w").write(subprocess.run([sys.executable, "tools/ncu_raw_summary.py", f"{G}prof_{tag}.ncu-rep"], capture_output=True, text=True).stdout)
raw = subprocess.run(["ncu", "-i", f"{G}prof_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); H, U = rr[0], rr[1]
col = H.index
def scaled(r, name, table):
    return float(r[col(name)].replace(",", "")) * table.get(U[col(name)], 1)
B = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}; T = {"ns": 1e-3, "us": 1, "ms": 1e3}
det = {}
for r in rr[2:]:
    key = [k for k in ("vertex", "geom", "clip", "large_bin", "order", "tile") if k in r[col("Kernel Name")]][0]
    rd, wr = scaled(r, "dram__bytes_read.sum", B), scaled(r, "dram__bytes_write.sum", B)
    det[key] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr, "duration_us": scaled(r, "gpu__time_duration.sum", T),
                "issue_active_pct": float(r[col("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
                "warp_inst": float(r[col("smsp__inst_executed.sum")].replace(",", "")), "regs": float(r[col("launch__registers_per_thread")])}
json.dump({"tile": det["tile"]["traffic_bytes"], "geometry": det["geom"]["traffic_bytes"] + det["vertex"]["traffic_bytes"],
           "_source": f"profiles/{RND}_ncu_full_summary.txt (ncu --set full --clock-control none --import-source on, one launch of each kernel of one C2 frame; tools/profile_round.sh)",
           "_commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or "see git log of profiles/traffic.json",
           "_detail": det}, open(P + "traffic.json", "w"), indent=1)
if TRAFFIC_ONLY:
    sys.exit(0)
shutil.copy(f"{G}tile_times_{tag}.txt", P + f"{RND}_tile_times.txt")
shutil.copy(f"{G}configs_{tag}.jsonl", P + f"{RND}_configs.jsonl")
open(P + f"{RND}_bench_n1.json", "w").write(open(f"{G}bench_{tag}.json").read().strip().splitlines()[-1] + "\n")
shutil.copy(f"{G}ref_{tag}.json", P + f"{RND}_bench_n1_reference_arm.json")
import os
if os.path.exists(f"{G}prof_{tag}_c3.ncu-rep"):
    c3 = subprocess.run([sys.executable, "tools/ncu_raw_summary.py", f"{G}prof_{tag}_c3.ncu-rep"], capture_output=True, text=True).stdout
    ph = subprocess.run([sys.executable, "tools/ncu_phase_split.py", f"{G}prof_{tag}_c3.ncu-rep"], capture_output=True, text=True).stdout
    open(P + f"{RND}_ncu_c3_tile_summary.txt", "w").write("# tile kernel of the C3 frame (250K near-clipped triangles, 3840x2160)\n" + c3 + "\n# warp instructions per phase (tools/ncu_phase_split.py)\n" + ph)
ph = subprocess.run([sys.executable, "tools/ncu_phase_split.py", f"{G}prof_{tag}.ncu-rep"], capture_output=True, text=True).stdout
open(P + f"{RND}_ncu_tile_phases.txt", "w").write("# C2 tile kernel: warp instructions and stall samples per phase (tools/ncu_phase_split.py)\n" + ph)
print("\n".join(out[-8:]))
for k, v in det.items(): print(k, round(v["duration_us"], 1), "us", round(v["warp_inst"] / 1e6, 2), "M inst", v["regs"], "regs", round(v["issue_active_pct"], 1), "% issue")
