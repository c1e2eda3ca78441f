for v in 0 1 2 1 2; do if [ $v = 0 ]; then python profiles/measure_configs.py 2>/dev/null | grep -A1 C3; else ICP_GPU_PROJ_HALF_TILES=$v python profiles/measure_configs.py 2>/dev/null | grep -A1 C3; fi; done
ICP_GPU_PROJ_HALF_TILES=1 python profiles/measure_sequence.py | grep "frames_per_s"
ICP_GPU_PROJ_HALF_TILES=2 python profiles/measure_sequence.py | grep "frames_per_s"
