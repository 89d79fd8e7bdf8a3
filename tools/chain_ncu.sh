#!/bin/bash
# ncu captures of the warp / blend chain -> gpurun_out/r02_chain_full_v2.ncu-rep (one frame, --set full, cold cache per replay) and
# gpurun_out/r02_chain_inpipe_v2.csv (8 frames, DRAM / L2 bytes with caches left alone, application replay).  Reduced by tools/chain_ncu_json.py.
K='regex:k_warp_rows|k_dt_|k_blur_blend|k_rowscan_bgrx'
ncu --set full --clock-control none --import-source on -k "$K" -s 32 -c 10 -o gpurun_out/r02_chain_full_v2 -f python tools/chain_only.py 8 > gpurun_out/r02_chain_full_v2.log 2>&1
ncu --cache-control none --clock-control none --replay-mode application --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum \
    -k "$K" --csv --log-file gpurun_out/r02_chain_inpipe_v2.csv python tools/chain_only.py 8 > gpurun_out/r02_chain_inpipe_v2.log 2>&1
tail -2 gpurun_out/r02_chain_full_v2.log; tail -2 gpurun_out/r02_chain_inpipe_v2.log
