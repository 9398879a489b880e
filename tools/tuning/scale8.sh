#!/bin/bash
# 8-GPU scaling on one box: the default line (peer-memory gather) + variants (brief)
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 $EXTRA > gpurun_out/r2f_n8_$tag.json 2> gpurun_out/r2f_n8_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2f_n8_$tag.json"))
    o=d.get("other_scaling") or {}
    print("$tag", "value %.3e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.3f"%d["e2e"]["ms_per_step"], "per-rank ms", [round(x,3) for x in d["per_rank_kernel_ms_per_step"]], "W", [round(x) for x in d["per_rank_power_w"]], "strong %.3f ms"%o.get("ms_per_step", float("nan")), "gather", (d.get("gather") or {}).get("mode"), (d.get("gather") or {}).get("verified"), o.get("gather_ok"))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r2f_n8_$tag.err").read()[-1500:])
PY
}
EXTRA="" run peer A=1
EXTRA="--brief --gather nccl" run nccl A=1
EXTRA="--brief --no-gather" run nogather A=1
