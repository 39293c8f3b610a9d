#!/bin/bash
# round 2, call 10: programmatic dependent launch A/B on the graphed UNet call (off / early trigger / late trigger)
mkdir -p gpurun_out
L=$PWD/emote_hack_b200/lib/libemote_b200_pdllate.so
echo "== PDL off";   EMOTE_PDL=0 timeout 300 python scripts/graph_unet.py 2>&1 | tail -5
echo "== PDL early"; EMOTE_PDL=1 timeout 300 python scripts/graph_unet.py 2>&1 | tail -5
echo "== PDL late";  EMOTE_PDL=1 EMOTE_B200_LIB=$L timeout 300 python scripts/graph_unet.py 2>&1 | tail -5
echo "== PDL off, late lib";  EMOTE_PDL=0 EMOTE_B200_LIB=$L timeout 300 python scripts/graph_unet.py 2>&1 | tail -4
echo "== tests with PDL late"
EMOTE_PDL=1 EMOTE_B200_LIB=$L timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -4
