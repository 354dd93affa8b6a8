#!/bin/bash
# A/B harness: build kernel variants (-D switches) into gpurun_out/variants/ HERE, then time each on the GPU box with
#   gpurun -- bash tools/ab_bench.sh run
# usage: tools/ab_bench.sh build name1 "-DFOO=1" name2 "-DBAR=2" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p ab_variants
if [ "$1" = "build" ]; then
  shift
  while [ $# -gt 1 ]; do
    name=$1; flags=$2; shift 2
    make -s -C meteoros_b200/csrc OUT=../../ab_variants/lib_$name.so EXTRA="$flags" -B > /dev/null
    grep -E "cloud_raymarch_kernelILb1ELb0ELb0" -A1 meteoros_b200/csrc/build.log | grep Used | sed "s/^/$name: /"
  done
  make -s -C meteoros_b200/csrc -B > /dev/null   # restore the default in-tree build
else
  for lib in ab_variants/lib_*.so; do
    echo "== $lib"
    METEOROS_B200_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], 'ms', d['value'], 'Mrays/s')"
  done
fi
