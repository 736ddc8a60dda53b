#!/bin/bash
# ncu --set full captures of the round's (f)-row kernels, each after its plain command exited 0:
#   bash tools/ncu_extra.sh <tag>   -> gpurun_out/<tag>_prefix_beam.ncu-rep, gpurun_out/<tag>_augment.ncu-rep
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python tools/prefix_beam_probe.py > $OUT/${TAG}_prefix_beam_plain.log 2>&1 || { echo "prefix probe failed"; exit 1; }
python tools/augment_probe.py > $OUT/${TAG}_augment_plain.log 2>&1 || { echo "augment probe failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:ctc_prefix_beam -s 16 -c 1 \
    -f -o $OUT/${TAG}_prefix_beam python tools/prefix_beam_probe.py > $OUT/${TAG}_ncu_pb.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:augment_lines -s 3 -c 1 \
    -f -o $OUT/${TAG}_augment python tools/augment_probe.py > $OUT/${TAG}_ncu_aug.log 2>&1
ls -la $OUT | tail -6
