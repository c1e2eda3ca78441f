B="python bench.py --steps 1 --warmup 3 --no-multi --no-cpu-baseline"
ICP_GPU_GROUP_MIN=8 ICP_GPU_MATCH_CHUNKS=1 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"knn_group" -s 40 -c 1 -f -o gpurun_out/prof_group $B > gpurun_out/ncu_group.log 2>&1
tail -3 gpurun_out/ncu_group.log
