"""Turns gpurun_out/*.ncu-rep + launches.csv into the tracked summaries under profiles/ and dumps
the SASS of every kernel the bench launches (cuobjdump -sass).  Run here (no GPU needed)."""
import csv, io, json, os, re, subprocess, sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
os.makedirs(OUT, exist_ok=True)

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


summary = {}
lines = [f"# ncu --set full summaries ({TAG}; B200, --clock-control none)\n"]
for name in sorted(os.listdir(GP)):
    if not (name.endswith(".ncu-rep") and name.startswith("prof_")):
        continue
    hdr, units, rows = raw(os.path.join(GP, name))
    ci = {h: i for i, h in enumerate(hdr)}
    workload = "config 5 (50 M triangles, 7.6 GB of nodes + triangles: the scene does not fit the 126 MB L2)" if "_c5_" in name else "config 3"
    for r in rows:
        k = r[ci["Kernel Name"]]
        lines.append(f"\n## {k} — {workload} ({name})\n\n| metric | value |\n|---|---|")
        d = {}
        for w in WANT:
            if w in ci:
                lines.append(f"| {w} | {r[ci[w]]} {units[ci[w]]} |")
                d[w] = r[ci[w]]
        stalls = [(h, float(r[i])) for h, i in ci.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        stalls.sort(key=lambda x: -x[1])
        lines.append("| top stalls (warps per issue) | " + ", ".join(f"{h[34:-24]} {v:.2f}" for h, v in stalls[:5]) + " |")
        summary.setdefault(k, []).append(d)
open(os.path.join(OUT, f"{TAG}_ncu_full.md"), "w").write("\n".join(lines) + "\n")

# launch list -> per-kernel share of one frame
lp = os.path.join(GP, "launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(l for l in open(lp) if not l.startswith("=="))]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    tot = defaultdict(float)
    cnt = Counter()
    for r in rows[1:]:
        if len(r) <= ci["Metric Value"]:
            continue
        k = re.sub(r"\(.*", "", r[ci["Kernel Name"]])
        k = re.sub(r"^void ", "", k)
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "nsecond": 1e-6, "msecond": 1.0}.get(unit, 1e-3)
        tot[k] += v
        cnt[k] += 1
    all_ms = sum(tot.values())
    with open(os.path.join(OUT, f"{TAG}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({TAG}): 2 config-3 frames, serialised, cold caches — compare SHARES\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"| {k} | {cnt[k]} | {v:.3f} | {100 * v / all_ms:.1f} % |\n")
    import shutil
    shutil.copy(lp, os.path.join(OUT, f"{TAG}_launches.csv"))

for j in ("bench.json", "bench_ref.json"):
    if os.path.exists(os.path.join(GP, j)):
        import shutil
        shutil.copy(os.path.join(GP, j), os.path.join(OUT, f"{TAG}_{j}"))

# counters for bench.py's roofline: warp instructions per ray and DRAM bytes per launch of the two traversal
# kernels, from the ncu --set full capture of the SAME command (scripts/gpu_profile.sh)
try:
    b = json.load(open(os.path.join(GP, "bench.json")))
    n_cam = 1921 * 1081 * 16                       # config 3: sampler extent x 16 spp
    n_sh = b["config"]["rays_per_frame"] - n_cam
    hdr, units, rows = raw(os.path.join(GP, "prof_k_trace.ncu-rep"))
    ci = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {"source": f"profiles/{TAG}_ncu_full.md (ncu --set full --clock-control none of scripts/prof_frame.py, one launch each)"}
    for r in rows:
        k = r[ci["Kernel Name"]]
        anyhit = "k_trace<(bool)1" in k or "k_trace<1" in k
        inst = float(r[ci["smsp__inst_executed.sum"]].replace(",", ""))
        dram = sum(float(r[ci[m]].replace(",", "")) * scale[units[ci[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        key = "any" if anyhit else "closest"
        out[f"{key}_warp_inst_per_ray"] = inst / (n_sh if anyhit else n_cam)
        out[f"{key}_warp_inst_per_launch"] = inst
        out[f"{key}_dram_bytes_per_launch"] = dram
        out[f"{key}_kernel"] = k
        out[f"{key}_issue_active_pct"] = float(r[ci["sm__issue_active.avg.pct_of_peak_sustained_elapsed"]])
        out[f"{key}_l1tex_data_pipe_pct"] = float(r[ci["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]])
        out[f"{key}_lanes_per_inst"] = float(r[ci["smsp__thread_inst_executed_per_inst_executed.ratio"]])
    cpath = os.path.join(OUT, f"{TAG}_counters.json")
    allc = json.load(open(cpath)) if os.path.exists(cpath) else {}
    allc["c3"] = out
    json.dump(allc, open(cpath, "w"), indent=1)
except Exception as e:
    print("counters:", e)

# SASS listings
sass_dir = os.path.join(OUT, "sass")
os.makedirs(sass_dir, exist_ok=True)
lib = os.path.join(ROOT, "pbrt_rust_b200", "libpbrtb200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
parts = re.split(r"\n\s*Function : ", txt)
keep = {"k_raygen_groups": "k_raygen_groups", "k_raygen_full": "k_raygen_full",
        "_Z7k_shadeILi8ELb0ELb0EE": "k_shade", "_Z7k_shadeILi8ELb1ELb0EE": "k_shade_image_textures",
        "_Z7k_shadeILi6ELb1ELb1EE": "k_shade_ext", "_Z12tex_eval_extILi3EE": "k_shade_ext_tex_eval_level3",
        "_Z12tex_eval_extILi0EE": "k_shade_ext_tex_eval_level0", "_Z6noise_fff": "k_shade_ext_noise",
        "k_halton_binILi0": "k_halton_bin_count", "k_halton_binILi1": "k_halton_bin_scatter",
        "k_halton_samples": "k_halton_samples",
        "_Z6k_film5DFilm": "k_film", "k_film_develop": "k_film_develop", "k_area_tri_setup": "k_area_tri_setup", "k_scatter_raster": "k_scatter_raster",
        "_Z6k_fold5DFold": "k_fold", "k_cost_probeILb0ELb0": "k_cost_probe_tri_single",
        "_Z7k_traceILb0ELb0ELb1ELi1ELi1ELi2EE": "k_trace_closest_tri_multi_camera_box2",
        "_Z7k_traceILb1ELb0ELb1ELi0ELi2ELi2EE": "k_trace_any_tri_multi_queue_box2",
        "_Z7k_traceILb1ELb0ELb0ELi0ELi2ELi2EE": "k_trace_any_tri_single_queue_box2",
        "_Z7k_traceILb0ELb0ELb0ELi1ELi1ELi2EE": "k_trace_closest_tri_single_camera_box2",
        "_Z7k_traceILb0ELb1ELb1ELi1ELi1ELi2EE": "k_trace_closest_spheres_multi_camera_box2",
        "_Z7k_traceILb0ELb0ELb0ELi0ELi1ELi2EE": "k_trace_closest_tri_single_buffer_box2",
        "_Z7k_traceILb0ELb0ELb0ELi1ELi1ELi3EE": "k_trace_closest_tri_single_camera_box3_ffma",
        "_Z7k_traceILb1ELb0ELb0ELi0ELi2ELi3EE": "k_trace_any_tri_single_queue_box3_ffma",
        "_Z7k_traceILb0ELb0ELb0ELi1ELi1ELi1EE": "k_trace_closest_tri_single_camera_box1"}
index = []
for p in parts[1:]:
    fn = p.split("\n", 1)[0].strip()
    for key, out in keep.items():
        if key in fn:
            body = "Function : " + p
            ops = Counter(m.group(1).split(".")[0] for m in re.finditer(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]+)", body))
            # drop the hex encodings (second column and encoding-only lines): mnemonics are the evidence
            lines = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln) for ln in body.split("\n")]
            body_txt = "\n".join(ln for ln in lines if ln.strip())
            open(os.path.join(sass_dir, out + ".sass"), "w").write(body_txt + "\n")
            index.append((out, fn, sum(ops.values()), ops.most_common(8)))
with open(os.path.join(sass_dir, "README.md"), "w") as f:
    f.write("# SASS listings (cuobjdump -sass libpbrtb200.so, sm_100a)\n\nNo tensor-core (UTC*MMA/HMMA) and no TMA "
            "instructions appear: nothing on this path is a dense contraction; node/triangle fetches are `LDG.E.128.CONSTANT`.\n\n")
    f.write("| listing | mangled name | SASS instructions | top opcodes |\n|---|---|---|---|\n")
    for out, fn, n, ops in sorted(index):
        f.write(f"| {out}.sass | `{fn}` | {n} | {', '.join(f'{o} {c}' for o, c in ops)} |\n")
print("profiles written")
