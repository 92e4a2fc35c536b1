#!/bin/bash
# time the global attention op with the default build and every variants/lib_*.so; timeline builds (lib_tl*) run twice only
echo "default:"; timeout 120 python tools/time_attention.py | grep "grid 64"
for f in variants/lib_*.so; do
  echo "$f:"
  case $f in
    *lib_tl*) ULLSAM_B200_LIB=$PWD/$f timeout 120 python tools/time_attention_once.py 2 | tail -4 ;;
    *) ULLSAM_B200_LIB=$PWD/$f timeout 120 python tools/time_attention.py | grep "grid 64" ;;
  esac
done
