#!/bin/bash
# radix-sort instances side by side (KSLAM_RS_CFG, radix_sort.cu): random 16-byte records, 8 passes over the full key
for cfg in ${CFGS:-0 10 9}; do
  for n in 32000000 128000000; do
    echo "== KSLAM_RS_CFG=$cfg N=$n"
    KSLAM_RS_CFG=$cfg N=$n REPS=4 python tools/prof_sort.py | tail -2
  done
done
for cfg in ${CFGS:-0 10 9}; do
  KSLAM_RS_CFG=$cfg python -m pytest tests/test_gpu_parity.py -q -k "radix or config1_shape or adversarial or golden" 2>&1 | tail -2
done
