#!/bin/bash
# per-rank phase times of the domain mode (BH_LET_TIMERS=1) at the given GPU counts
for np in "$@"; do
  BH_LET=1 BH_LET_MIN_WORLD=2 BH_LET_TIMERS=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29800+np)) bench.py --gpus $np --steps 16 --warmup 3 > gpurun_out/lett_$np.json 2> gpurun_out/lett_$np.err
  python - <<PY
import json
txt=open("gpurun_out/lett_$np.json").read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{"metric')][0])
print("N=$np ms/step", round(d["ms_per_step"],3), d["phases_ms_per_evaluation"])
PY
  grep -h "let rank" gpurun_out/lett_$np.err | sort -u | head -20
done
