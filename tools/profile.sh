#!/bin/bash
# One-command profile capture for a round (run under gpurun on ONE GPU):
#   gpurun --timeout 600 -- 'bash tools/profile.sh v17'
# writes gpurun_out/launches_<tag>.csv (every launch with its device time, DRAM bytes and warp instructions — cold-cache and
# serialised: compare shares, not absolutes), gpurun_out/raster_<tag>.ncu-rep (ncu --set full of cpvk_k_raster on C3/M1) and
# gpurun_out/raster_c4_<tag>.ncu-rep (the same kernel on a C4-shaped draw). Back in the build container:
#   python tools/ncu_summary.py gpurun_out/raster_<tag>.ncu-rep > profiles/raster_r<round>_<tag>_ncu.txt
#   python tools/ncu_sass_hot.py gpurun_out/raster_<tag>.ncu-rep            # executed-weighted SASS hot spots
# Numbers printed by a run under ncu are never bench values.
set -u
tag=${1:-run}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cpvk_k_raster -s 3 -c 1 -f -o gpurun_out/raster_$tag \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_raster_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cpvk_k_raster -s 1 -c 1 -f -o gpurun_out/raster_c4_$tag \
    python tools/probe_overdraw_small.py > gpurun_out/ncu_raster_c4_$tag.log 2>&1
ls -la gpurun_out/*_$tag.*
