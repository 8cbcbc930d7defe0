#!/bin/bash
# SASS listings of the hot kernels of the in-tree build, for profiles/sass/ (evidence for UTMALDG, VABSDIFF4 / VIMNMX.U16x2, SYNCS, LDGSTS, IDP).
TAG=${1:-r02}
mkdir -p profiles/sass
for k in k_inter_tma k_inter_list k_deblock3 k_bs_prog2 k_intra k_residual; do
  cuobjdump -sass h264_video_decoder_demo_b200/libh264b2.so | awk -v K="$k" '/Function : /{f = index($0, K) > 0} f' | sed -E 's/ +\/\* 0x[0-9a-f]+ \*\/$//' | grep -v "^\s*$" > profiles/sass/${TAG}_$k.sass
done
grep -c UTMALDG profiles/sass/${TAG}_k_inter_tma.sass
