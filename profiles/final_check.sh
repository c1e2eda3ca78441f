# Round-2 last session: full GPU test suite, smoke(), and the bench line at HEAD (cpu baseline legs skipped here: the driver runs them)
timeout 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/final_tests.log; cat gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
python - <<'P'
import json
d = json.load(open('gpurun_out/bench_final_n1.json'))
print('ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'launches', d['gpu_launches_per_step'], d['stage_ms_per_iteration'], 'pq44', round(d['pair_queue_44']['pairs_per_s'], 1), 'sharded single', round(d['sharded_3m']['ms_single_gpu'], 2), 'chk', d['pose_checksum'], d['clocks'])
P
