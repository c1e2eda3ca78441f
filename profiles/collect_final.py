"""Copies the outputs of profiles/final_measure.sh (gpurun_out/final/) into profiles/ as r2_* and regenerates the summaries derived
from them: launch-list summaries (hot and cold caches), ncu metric excerpts of the iteration / build / projective kernels, the DRAM
traffic per launch that bench.py reads, SASS listings of every kernel of lib/libicp_gpu.so and the ptxas resource log.
Usage (in the build container, after gpurun merged gpurun_out/final): python profiles/collect_final.py"""
import csv, glob, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = os.path.join(ROOT, "gpurun_out", "final"); P = os.path.join(ROOT, "profiles")
R = "r2_"


def copy(src, dst):
    if os.path.exists(os.path.join(F, src)):
        shutil.copy(os.path.join(F, src), os.path.join(P, R + dst))
    else:
        print("missing:", src)


for src, dst in [("bench_n1.json", "bench_n1.json"), ("bench_n2.json", "bench_n2.json"), ("bench_n4.json", "bench_n4.json"), ("bench_n8.json", "bench_n8.json"),
                 ("bench_reference_arm.json", "bench_reference_arm.json"), ("config_timings.json", "config_timings.json"),
                 ("build_timings.json", "build_timings.json"), ("reduce_profile.json", "reduce_profile.json"), ("reduce_profile_3M.json", "reduce_profile_3M.json"),
                 ("sequence.json", "sequence_tum_shaped.json"), ("sequence_frame_breakdown.json", "sequence_frame_breakdown.json"),
                 ("normals_depth.json", "normals_depth_timings.json"), ("pair_queue_contexts.txt", "pair_queue_contexts_per_gpu.txt"),
                 ("timeline_1chunk.json", "timeline_one_iteration_1chunk.json"), ("timeline_2chunks.json", "timeline_one_iteration_2chunks.json"),
                 ("sharded_detail_n2.json", "sharded_detail_n2.json"), ("launches.csv", "launches_cold_caches.csv"), ("launches_hot.csv", "launches_hot_caches.csv")]:
    copy(src, dst)
for tag in ("cold", "hot"):
    lst = os.path.join(P, f"{R}launches_{tag}_caches.csv")
    if os.path.exists(lst):
        with open(os.path.join(P, f"{R}launches_{tag}_caches_summary.txt"), "w") as f:
            subprocess.run([sys.executable, os.path.join(P, "summarize_launches.py"), lst], stdout=f, check=True)
for rep, out in (("prof_hot.ncu-rep", "ncu_full_hot_kernels.txt"), ("prof_build.ncu-rep", "ncu_full_build_kernels.txt"), ("prof_proj.ncu-rep", "ncu_full_projective_kernel.txt")):
    if os.path.exists(os.path.join(F, rep)):
        with open(os.path.join(P, R + out), "w") as f:
            subprocess.run([sys.executable, os.path.join(P, "ncu_metrics.py"), os.path.join(F, rep)], stdout=f, check=True)
rep = os.path.join(F, "prof_hot.ncu-rep")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines())); h, u = rows[0], rows[1]


def val(d, k):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u[h.index(k)], 1)
    return float(d[k].replace(",", "")) * mult


ks = {}
for r in rows[2:]:
    d = dict(zip(h, r))
    name = d["Kernel Name"].replace("(bool)", "").replace("(int)", "")
    ks[name] = {"dram_bytes_read": val(d, "dram__bytes_read.sum"), "dram_bytes_write": val(d, "dram__bytes_write.sum"), "gpu_time_us": val(d, "gpu__time_duration.sum")}
json.dump({"source": "ncu --set full --clock-control none, one steady-state iteration of `bench.py --steps 2 --warmup 3` (profiles/final_measure.sh); per launch; "
                     "caches cold under ncu replay; with two chunk chains a launch of knn_prep / knn_bvh covers half of the queries",
           "kernels": ks}, open(os.path.join(P, R + "ncu_traffic.json"), "w"), indent=1)

# SASS of every kernel (cuobjdump of the library that was measured), one file per source file, and the ptxas resource lines
lib = os.path.join(ROOT, "icp_variants_b200", "lib", "libicp_gpu.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
# keep address + instruction, drop the 128-bit encodings (two hex comments per instruction): a third of the size, nothing a reader uses
sass = re.sub(r"[ \t]*/\* 0x[0-9a-f]{16} \*/", "", sass)
sass = "\n".join(ln.rstrip() for ln in sass.splitlines() if ln.strip()) + "\n"
funcs = re.split(r"(?=\n\s*Function : )", sass)
groups = {"knn_prep": "knn_prep_kernel", "knn_group": "knn_group_kernel", "knn_bvh": "knn_bvh_kernel", "reduce": "reduce_kernel", "projective": "projective_kernel", "lm_eval": "lm_eval_kernel",
          "pca_normals": "pca_normals_kernel", "depth_cloud": "depth_cloud|flag_count|block_scan|flag_scatter",
          "index_build": "pack_bbox|pack_normals|keys_kernel|radix_|gather_records|level_flags|level_rank|upper_levels|bvh_level|leaf_adjacency|seed_from_keys",
          "other": None}
used = set()
for g, pat in groups.items():
    sel = []
    for i, fn in enumerate(funcs):
        m = re.search(r"Function : (\S+)", fn)
        if not m:
            continue
        if (pat is not None and re.search(pat, m.group(1))) or (pat is None and i not in used):
            sel.append(fn); used.add(i)
    if sel:
        # first template instance only for the big kernels, to keep the listings reviewable; every instance is named in the header
        names = [re.search(r"Function : (\S+)", fn).group(1) for fn in sel]
        prefer = {"reduce": "ILi1ELb1E", "knn_prep": "ILb0ELb0E", "knn_group": "ILb0E", "knn_bvh": "ILb0ELb0E", "projective": "ILi256E"}.get(g)    # the instance the bench runs
        first = [fn for fn in sel if prefer and prefer in re.search(r"Function : (\S+)", fn).group(1)][:1] or sel[:1]
        keep = sel if g in ("index_build", "depth_cloud", "other") else first
        with open(os.path.join(P, f"{R}sass_{g}.txt"), "w") as f:
            f.write("cuobjdump -sass icp_variants_b200/lib/libicp_gpu.so (sm_100a); instances in the library: " + ", ".join(names) + "\n")
            f.write("".join(keep))
with open(os.path.join(P, R + "ptxas_v.txt"), "w") as f:
    for log in sorted(glob.glob(os.path.join(ROOT, "icp_variants_b200", "lib", "*.ptxas.log"))):
        f.write("== " + os.path.basename(log) + "\n")
        lines = open(log).read().splitlines()
        for i, ln in enumerate(lines):
            if "Compiling entry function" in ln:
                f.write(ln.split("'")[1] + "\n")
            elif "Used " in ln or "spill" in ln:
                f.write("    " + ln.strip() + "\n")
print(json.dumps(ks, indent=1))
