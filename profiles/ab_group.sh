# A/B of the group search (ICP_GPU_GROUP_MIN=0 off / 8 on): the bench line (headline, pair_queue_44, sharded_3m on one GPU) and the config timings
for g in 0 8; do
  ICP_GPU_GROUP_MIN=$g python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gm$g.json 2> gpurun_out/bench_gm$g.err
  ICP_GPU_GROUP_MIN=$g python profiles/measure_configs.py > gpurun_out/configs_gm$g.json 2> gpurun_out/configs_gm$g.err
done
python - <<'P'
import json
for g in (0, 8):
    d = json.load(open(f'gpurun_out/bench_gm{g}.json'))
    print('gm', g, 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'launches', d['gpu_launches_per_step'], 'pq44', round(d['pair_queue_44']['pairs_per_s'], 1), round(d['pair_queue_44']['ms_total'], 1), 'sharded single', round(d['sharded_3m']['ms_single_gpu'], 2), 'chk', d['pose_checksum'])
    c = json.load(open(f'gpurun_out/configs_gm{g}.json'))
    print({k: round(v['ms'], 3) for k, v in c.items() if isinstance(v, dict) and 'ms' in v})
P
