#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus 2): the engine's device-side slab gather against the single-GPU result and
# against the NCCL stage-level path; bench in replica mode (weak scaling) and shard mode (one volume over the ranks).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
NG=${1:-2}
nvidia-smi -L | tee $O/multi_gpus.txt
nvidia-smi topo -m > $O/multi_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -6 $O/pytest_multi.log
run() { # name, extra bench args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG $2 > $O/multi_$1.json 2> $O/multi_$1.err; echo "$1 rc=$?"
}
run replica_C2 "--steps 4 --warmup 3"
run shard_C2 "--mode shard --steps 6 --warmup 3"
run shard_C3 "--mode shard --config C3 --steps 4 --warmup 3"
timeout 400 python bench.py --mode shard --config C3 --steps 4 --warmup 3 > $O/multi_shard_C3_n1.json 2> $O/multi_shard_C3_n1.err; echo "shard n1 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/multi_*.json")):
    try:
        d = json.load(open(f))
        print(f, "n", d["n_gpus"], d["scaling"], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), d["config"]["stage_ms_per_step"], "dice", d.get("dice"))
    except Exception as e:
        print(f, "unreadable:", e, open(f.replace(".json", ".err")).read()[-600:])
PY
