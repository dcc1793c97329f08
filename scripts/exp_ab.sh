#!/bin/bash
# A/B of two builds of the library on the same box: liblwpose_b200_base.so (A) against liblwpose_b200.so (B)
P="lightweight-human-pose-estimation.pytorch_b200"
cp $P/liblwpose_b200.so /tmp/new.so
for r in 1 2; do
  cp $P/liblwpose_b200_base.so $P/liblwpose_b200.so; echo "A (base)"; python scripts/time_layers.py "$@"
  cp /tmp/new.so $P/liblwpose_b200.so; echo "B (new)"; python scripts/time_layers.py "$@"
done
