for g in ${GMS:-8}; do echo "GROUP_MIN=$g"; ICP_GPU_GROUP_MIN=$g python bench.py --steps 10 --warmup 3 --no-multi --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms_per_iteration'], d['pose_checksum'], d['gpu_launches_per_step'])"; done
GROUP_MINS=${GMS:-8} python profiles/probe_group.py | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items():
    for g,r in v.items(): print('pair',k,'group_min',g,'ms %.3f'%r['ms_30_iterations'],'group+walk us %.1f'%r['group_and_walk_us'],'chk',r['pose_checksum'],{a:round(b,1) for a,b in r['per_iteration'].items()})"
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_size_eth or chunk_chains or config4_full or index_edge or bench_config" 2>&1 | tail -3
B="python bench.py --steps 1 --warmup 3 --no-multi --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -k regex:"knn_|reduce_kernel" -s 700 -c 130 --csv --log-file gpurun_out/group_launches.csv $B > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/group_launches.csv 2>&1 | tail -5
