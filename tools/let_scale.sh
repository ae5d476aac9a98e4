#!/bin/bash
# domain mode vs replicated tree on one box: bench.py at N GPUs, 1M bodies per GPU (weak scaling)
out=${OUT:-gpurun_out/letscale}
run() { # name, nproc, env...
  name=$1; np=$2; shift 2
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29700+np)) bench.py --gpus $np --steps 16 --warmup 3 > ${out}_$name.json 2> ${out}_$name.err
  python - <<PY
import json
txt=open("${out}_$name.json").read()
try:
    d=json.loads([l for l in txt.splitlines() if l.startswith('{"metric')][0])
    s=d.get("domain_mode_rank0") or {}
    print("$name", "ms/step", round(d["ms_per_step"],3), "steps/s", round(d["steps_per_s"],1), "inter/s %.3e" % d["value"], d["phases_ms_per_evaluation"], "LET cells", s.get("let_cells"), "imported", s.get("cells_imported"), "fallbacks", s.get("fallbacks"), "mode", s.get("enabled"))
except Exception as ex:
    print("$name failed", ex, txt[-300:])
PY
  grep -h "let rank" ${out}_$name.err | sort | head -8
}
for np in "$@"; do
  run let$np $np BH_LET=1 BH_LET_MIN_WORLD=2
  run repl$np $np BH_LET=0
done


