#!/bin/bash
# round-2 ncu evidence: launch list of the default bench command, --set full of the dominant kernels (never a bench value)
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --side-configs none --steps 1 --warmup 3"
# 1. every launch of two steps of the C3 sweep past the session start (serialised, cold cache: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 3200 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_launches.log 2>&1
# 2. the tile kernel and the brush kernel of a 50 %-radius dab (the 200th dab of a step), full set
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_normals_tile|k_brush" -s 400 -c 2 -f -o gpurun_out/r2_c3_tile_brush $B > gpurun_out/r2_full_c3.log 2>&1
# 3. C2: the smooth brush's averaging kernel (L2 hit rate of the neighbour reads)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_smooth_a" -s 300 -c 2 -f -o gpurun_out/r2_c2_smooth python bench.py --config c2 --no-cpu-baseline --side-configs none --steps 1 --warmup 3 > gpurun_out/r2_full_c2.log 2>&1
ls -la gpurun_out/r2_*
