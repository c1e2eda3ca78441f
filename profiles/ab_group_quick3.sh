python profiles/measure_configs.py 2>/dev/null | python -c "
import json,sys
c=json.load(sys.stdin); print({k: round(v['ms'], 3) for k, v in c.items() if isinstance(v, dict) and 'ms' in v})"
PAIRS=0,28,34 GROUP_MINS="8" python profiles/probe_group.py | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items():
    for g,r in v.items(): print('pair',k,'group_min',g,'ms %.3f'%r['ms_30_iterations'],'group+walk us %.1f'%r['group_and_walk_us'],{a:round(b,1) for a,b in r['per_iteration'].items() if a in ('groups_started','groups_finished','members_finished','leaf_scans_per_group')})"
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_size_eth or chunk_chains or index_edge or bench_config" 2>&1 | tail -3
