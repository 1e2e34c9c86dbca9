#!/bin/bash
# the built library and every variant library under tests/probes/_variants through rank_variants.py
TAG=shipped python tests/probes/rank_variants.py "$@"
for lib in tests/probes/_variants/lib_*.so; do
  [ -e "$lib" ] || continue
  TAG="$(basename $lib .so)" DALIID_B200_LIB=$PWD/$lib python tests/probes/rank_variants.py "$@"
done
TAG=shipped python tests/probes/rank_variants.py "$@"
