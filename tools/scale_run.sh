#!/bin/bash
# weak-scaling sweep of bench.py on one box: N = 1, 2, 4, 8 (what the driver's SCALE run does)
out=${1:-gpurun_out/scale}
mkdir -p $(dirname $out)
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > ${out}_n$n.json 2> ${out}_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 > ${out}_n$n.json 2> ${out}_n$n.err
  fi
  python - <<PY
import json
try:
    d = json.load(open("${out}_n$n.json"))
    print($n, "ms/step", round(d["ms_per_step"], 3), "value %.3e" % d["value"], d["phases_ms_per_evaluation"], "e2e", round(d["e2e"]["steps_per_s"], 1))
except Exception as ex:
    print($n, "failed", ex)
PY
done
