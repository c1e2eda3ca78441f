python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { echo "== $*"; env $* python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.readline())
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms_per_iteration']['match_tree_walk'], d['pose_checksum'], d['gpu_launches_per_step'])
print(d['pair_queue_44']['pairs_per_s'], d['sharded_3m'].get('ms_single_gpu'))"; }
run ICP_GPU_MATCH_CHUNKS=1
run A=1
run ICP_GPU_MATCH_CHUNKS=1
run A=1
python profiles/measure_configs.py
