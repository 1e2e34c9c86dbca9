#!/bin/bash
# the built library (and every variant library under tests/probes/_variants) through rank_variants.py
python -m pytest tests/test_gpu_rank.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -2
DALI_RANK_V3=0 TAG=v2_control python tests/probes/rank_variants.py "$@"
for cfg in "DALI_RANK_V3_THREADS=256" "DALI_RANK_V3_THREADS=128"; do
  env $cfg TAG="main_$(echo $cfg | sed 's/DALI_RANK_V3_//g; s/ /_/g')" python tests/probes/rank_variants.py "$@"
done
for lib in tests/probes/_variants/lib_*.so; do
  [ -e "$lib" ] || continue
  for cfg in "DALI_RANK_V3_THREADS=256" "DALI_RANK_V3_THREADS=128"; do
    env $cfg TAG="$(basename $lib .so)_$(echo $cfg | sed 's/DALI_RANK_V3_//g; s/ /_/g')" DALIID_B200_LIB=$PWD/$lib python tests/probes/rank_variants.py "$@"
  done
done
