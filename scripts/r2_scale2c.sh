#!/bin/bash
mkdir -p gpurun_out
run() {
  label=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 12 --warmup 5 "$@" 2>/dev/null | grep -E "^\{" > gpurun_out/r2_scale2c_$label.json
  python - "$label" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r2_scale2c_{sys.argv[1]}.json").read())
print(sys.argv[1], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "bucket", d["config"].get("ddp_bucket_mb"))
PY
}
{
  timeout 300 python bench.py --steps 12 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-config1 2>/dev/null | grep -E "^\{" > gpurun_out/r2_scale2c_n1.json
  python -c "
import json; d=json.loads(open('gpurun_out/r2_scale2c_n1.json').read()); print('n1 value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"
  run b25 --bucket-mb 25
  run b1024 --bucket-mb 1024
  run b1024_syncbn --bucket-mb 1024 --sync-bn
} 2>&1 | tee gpurun_out/r2_scale2c.txt
