#!/bin/bash
# frame times (tools/frame_time.py) for every ab_variants/lib_*.so
cd "$(dirname "$0")/.."
for lib in ab_variants/lib_*.so; do
  echo "== $lib"
  METEOROS_B200_LIB=$PWD/$lib python tools/frame_time.py
  METEOROS_B200_LIB=$PWD/$lib python tools/frame_time.py --txaa
  METEOROS_B200_LIB=$PWD/$lib python tools/frame_time.py --txaa --no-godrays
  METEOROS_B200_LIB=$PWD/$lib python tools/frame_time.py --width 3840 --height 2160 --txaa
done
