# Group search per pair (0 = the bench pair, 28 / 34 = far-heavy pairs of the 44-pair sequence), off / on, with the diagnostic counters when
# the groupprobe build is selected (ICP_GPU_LIB_NAME=libicp_gpu_groupprobe.so), then the parity tests of the search
run() { python profiles/probe_group.py | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items():
    for g,r in v.items(): print('pair',k,'group_min',g,'ms %.3f'%r['ms_30_iterations'],'group+walk us %.1f'%r['group_and_walk_us'],'chk',r['pose_checksum'],{a:round(b,1) for a,b in r['per_iteration'].items() if a in ('groups_started','groups_finished','members_finished','leaf_scans_per_group','cycles_longest_group')})"; }
PAIRS=${PAIRS:-0,34,28} GROUP_MINS="${GMS:-0 8}" run
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_size_eth or chunk_chains or config4_full or index_edge or bench_config" 2>&1 | tail -3
