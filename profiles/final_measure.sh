# Round-1 final measurements on one B200 (run under gpurun from the repo root); outputs under gpurun_out/final/.
mkdir -p gpurun_out/final
python bench.py > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/final/bench_reference_arm.json 2> gpurun_out/final/bench_reference_arm.err
python profiles/measure_configs.py > gpurun_out/final/config_timings.json 2> gpurun_out/final/config_timings.err
python profiles/measure_reduce_profile.py > gpurun_out/final/reduce_profile.json 2> gpurun_out/final/reduce_profile.err
# launch list of the same bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final/ncu_launches.log 2>&1
# one full capture of the hot kernels (steady-state iteration: skip the first launches)
ncu --set full --clock-control none --import-source on -k regex:"knn_prep|knn_bvh|reduce_kernel" -s 60 -c 3 -f -o gpurun_out/final/prof_hot python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final/ncu_full.log 2>&1
tail -2 gpurun_out/final/bench_n1.json | cut -c1-300
python profiles/measure_sequence.py > gpurun_out/final/sequence.json 2> gpurun_out/final/sequence.err
python profiles/measure_normals.py > gpurun_out/final/normals_depth.json 2> gpurun_out/final/normals_depth.err
