"""Copies the outputs of profiles/final_measure.sh (gpurun_out/final/) into profiles/ and regenerates the summaries derived from
them (launch-list summary, ncu metric excerpt, DRAM traffic per launch read by bench.py)."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = os.path.join(ROOT, "gpurun_out", "final"); P = os.path.join(ROOT, "profiles")
for src, dst in [("bench_n1.json", "r1_bench_n1.json"), ("bench_reference_arm.json", "r1_bench_reference_arm.json"), ("config_timings.json", "r1_config_timings.json"),
                 ("reduce_profile.json", "r1_reduce_profile.json"), ("sequence.json", "r1_sequence_tum_shaped.json"), ("normals_depth.json", "r1_normals_depth_timings.json"),
                 ("launches.csv", "r1_launches.csv")]:
    shutil.copy(os.path.join(F, src), os.path.join(P, dst))
with open(os.path.join(P, "r1_launches_summary.txt"), "w") as f:
    subprocess.run([sys.executable, os.path.join(P, "summarize_launches.py"), os.path.join(P, "r1_launches.csv")], stdout=f, check=True)
rep = os.path.join(F, "prof_hot.ncu-rep")
with open(os.path.join(P, "r1_ncu_full_hot_kernels.txt"), "w") as f:
    subprocess.run([sys.executable, os.path.join(P, "ncu_metrics.py"), rep], stdout=f, check=True)
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
json.dump({"source": "ncu --set full --clock-control none, one steady-state iteration of `bench.py --steps 2 --warmup 1` (profiles/final_measure.sh); per launch; caches cold under ncu replay",
           "kernels": ks}, open(os.path.join(P, "r1_ncu_traffic.json"), "w"), indent=1)
print(json.dumps(ks, indent=1))
