#!/bin/bash
# ncu --set full capture of the MC and deblocking kernels over the FIRST pass of a 10-picture run (128 streams), for profiles/traffic_r02.json.
# The report is turned into its raw-page CSV on the box (the .ncu-rep is > 100 MB and gpurun brings back 64 MB at most);
# tools/make_traffic.py turns the CSV into per-kernel DRAM bytes next to the algorithmic bytes of the captured pictures.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:"k_inter_tma|k_inter_list|k_deblock3|k_bs_prog2" -c 36 -f -o /tmp/prof_$TAG python bench.py --streams 128 --max-pictures 10 --steps 1 --warmup 1 --no-e2e --no-cpu --no-bitstream --no-configs > gpurun_out/ncu_full_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_$TAG.raw.csv 2>/dev/null
python -c "from h264_video_decoder_demo_b200 import build as b; print(b.source_hash())" > gpurun_out/prof_$TAG.libsha1
ls -la gpurun_out/ | tail -5
