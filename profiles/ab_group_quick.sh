run() { PAIRS=${PAIRS:-0} GROUP_MINS=${GMS:-8} python profiles/probe_group.py | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items():
    for g,r in v.items(): print('pair',k,'group_min',g,'ms %.3f'%r['ms_30_iterations'],'group+walk us %.1f'%r['group_and_walk_us'],'chk',r['pose_checksum'],{a:round(b,1) for a,b in r['per_iteration'].items()})"; }
echo default; run
for v in ${VARIANTS:-B C D E}; do echo variant $v; ICP_GPU_LIB_NAME=libicp_gpu_v$v.so run; done
