#!/bin/bash
mkdir -p gpurun_out
run() {
  n=$1; label=$2; shift; shift
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $n --steps 12 --warmup 5 "$@" 2>/dev/null | grep -E "^\{" > gpurun_out/r2_scale8b_$label.json
  python - "$label" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r2_scale8b_{sys.argv[1]}.json").read())
print(sys.argv[1], "n", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "bucket", d["config"].get("ddp_bucket_mb"), "sync_bn", d["config"]["sync_bn"])
PY
}
{
  run 8 n8_b25 --bucket-mb 25
  run 8 n8_b1024 --bucket-mb 1024
  run 8 n8_b1024_samedata --bucket-mb 1024 --same-data
} 2>&1 | tee gpurun_out/r2_scale8b.txt
