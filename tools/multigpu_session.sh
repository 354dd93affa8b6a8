#!/bin/bash
# One multi-GPU measurement session (run under gpurun --gpus N): the driver's own N>1 command (views + sharded_8k sub-record),
# the reference arm under torchrun, the sharded 8K frame with every gather mode, and the bit-identity test.
#   gpurun --gpus N -- bash tools/multigpu_session.sh N [tag]
N=${1:-2}; TAG=${2:-r2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29501 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_n${N}_default.json 2> gpurun_out/${TAG}_n${N}_default.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_n${N}_default.json"))
    print("default N=$N: views", d["value"], "Mrays/s", d["ms_per_step"], "ms; e2e", d["e2e"]["value"])
    print("  sharded_8k:", json.dumps(d.get("sharded_8k")))
    print("  sharded_8k_f16:", json.dumps(d.get("sharded_8k_f16")))
    print("  e2e_f16:", json.dumps(d.get("e2e_f16")))
except Exception as e:
    print("default run failed:", e); print(open("gpurun_out/${TAG}_n${N}_default.err").read()[-1500:])
PY
run8k() {  # name, extra args
  NAME=$1; shift
  $TR --master-port 29502 bench.py --gpus $N --steps 10 --warmup 3 --workload frame8k "$@" > gpurun_out/${TAG}_n${N}_8k_$NAME.json 2> gpurun_out/${TAG}_n${N}_8k_$NAME.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_n${N}_8k_$NAME.json"))
    print("frame8k N=$N $NAME:", d["ms_per_step"], "ms", "rank_ms", d.get("rank_ms"), "clock samples", d["clocks"].get("samples"))
except Exception as e:
    print("frame8k $NAME failed:", e); print(open("gpurun_out/${TAG}_n${N}_8k_$NAME.err").read()[-800:])
PY
}
if [ "$LEAN" = "2" ]; then exit 0; fi
if [ "$LEAN" = "1" ]; then
  run8k local --gather local
  run8k peer_rows32 --gather peer_store --tile-rows 32
  exit 0
fi
run8k peer_store --gather peer_store
run8k bulk_store --gather bulk_store
run8k local --gather local
run8k peer_topdown --gather peer_store --ctx-flags 8
run8k peer_rows32 --gather peer_store --tile-rows 32
if [ "$FORWARD" = "1" ]; then run8k forward --gather forward; fi
OMP_NUM_THREADS=1 $TR --master-port 29503 bench.py --gpus $N --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_n${N}_ref.json 2> gpurun_out/${TAG}_n${N}_ref.err
python -c "import json; d=json.load(open('gpurun_out/${TAG}_n${N}_ref.json')); print('reference arm under torchrun:', d['value'], 'Mrays/s, cores', d['cpu_baseline']['cores'])"
python -m pytest tests/test_gpu_parity.py -m gpu -q -k multi_gpu 2>&1 | tail -2
