#!/bin/bash
mkdir -p gpurun_out
run() {
  label=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 4 "$@" 2>/dev/null | grep -E "^\{" > gpurun_out/r2_scale2_$label.json
  python - "$label" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r2_scale2_{sys.argv[1]}.json").read())
print(sys.argv[1], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "sync_bn", d["config"]["sync_bn"])
PY
}
{
  timeout 300 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-ref-cuda --no-config1 2>/dev/null | grep -E "^\{" > gpurun_out/r2_scale2b_n1.json
  python -c "
import json; d=json.loads(open('gpurun_out/r2_scale2b_n1.json').read()); print('n1 value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2))"
  run plain
  run samedata --same-data
  run syncbn_fused --sync-bn
  run rpvnet --config rpvnet34
} 2>&1 | tee gpurun_out/r2_scale2b.txt
