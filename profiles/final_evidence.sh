# Round-2 last session: GPU tests at HEAD, hot-cache launch list of the bench command and one full capture of the search kernels (one chain)
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 > gpurun_out/final_tests2.log; cat gpurun_out/final_tests2.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-multi"
$B > gpurun_out/plain_final.log 2>&1 && ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 900 --csv --log-file gpurun_out/launches_hot_final.csv $B > gpurun_out/ncu_launches_final.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_hot_final.csv > gpurun_out/launches_hot_final_summary.txt 2>&1; head -8 gpurun_out/launches_hot_final_summary.txt
export ICP_GPU_MATCH_CHUNKS=1
$B > gpurun_out/plain_final.log 2>&1 && ncu --set full --clock-control none --cache-control none --import-source on -k regex:"knn_prep|knn_group|knn_bvh" -s 90 -c 3 -f -o gpurun_out/prof_search_final $B > gpurun_out/ncu_full_final.log 2>&1
tail -2 gpurun_out/ncu_full_final.log | cut -c1-200
