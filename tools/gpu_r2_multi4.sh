#!/bin/bash
# Round 2, 4-GPU call (gpurun --gpus 4): the engine's slab gather with four ranks (bit-identity test) and the shard-mode
# bench lines at N = 4 (strong scaling of ONE volume; the replica scaling run is the driver's).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
NG=${1:-4}
nvidia-smi -L | tee $O/multi4_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > $O/pytest_multi4.log 2>&1; echo "pytest multi rc=$?"; tail -4 $O/pytest_multi4.log
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG $2 > $O/multi4_$1.json 2> $O/multi4_$1.err; echo "$1 rc=$?"; }
run shard_C2 "--mode shard --steps 6 --warmup 3"
run shard_C3 "--mode shard --config C3 --steps 4 --warmup 3"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/multi4_*.json")):
    try:
        d = json.loads([l for l in open(f).read().split("\n") if l.startswith("{")][-1])
        print(f, "n", d["n_gpus"], d["scaling"], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), d["config"]["stage_ms_per_step"], "dice", d.get("dice"))
    except Exception as e:
        print(f, "unreadable:", e, open(f.replace(".json", ".err")).read()[-600:])
PY
