# usage: bash profiles/sweep_knobs.sh "ENV1=.." "ENV2=.." ...   (one bench.py run per argument; prints the key numbers)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "== $*"; env $* python bench.py 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.readline())
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms_per_iteration'], d['roofline_fp32']['distance_evals_per_launch'], d['roofline_fp32']['nodes_per_launch'], d['pose_checksum'])"; }
for a in "$@"; do run $a; done
