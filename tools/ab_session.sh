#!/bin/bash
# One gpurun call: GPU test suite on the in-tree build, then the 4K bench for every ab_variants/lib_*.so and the warm 1080p pass
# times for every ab_variants/post/lib_*.so.  usage (on the GPU box): bash tools/ab_session.sh <tag> [notest]
cd "$(dirname "$0")/.."
tag=${1:-ab}
mkdir -p gpurun_out
if [ "$2" != "notest" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
  tail -3 gpurun_out/${tag}_pytest.log
fi
for lib in ab_variants/lib_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib" | tee -a gpurun_out/${tag}_bench.log
  METEOROS_B200_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], 'ms', d['value'], 'Mrays/s')" | tee -a gpurun_out/${tag}_bench.log
done
for lib in ab_variants/post/lib_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib" | tee -a gpurun_out/${tag}_passes.log
  METEOROS_B200_LIB=$PWD/$lib python tools/pass_times.py 2>&1 | tee -a gpurun_out/${tag}_passes.log
done
