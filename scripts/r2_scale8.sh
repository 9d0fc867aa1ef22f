#!/bin/bash
mkdir -p gpurun_out
run() {
  n=$1; label=$2; shift; shift
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $n --steps 12 --warmup 5 "$@" 2>/dev/null | grep -E "^\{" > gpurun_out/r2_scale8_$label.json
  python - "$label" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r2_scale8_{sys.argv[1]}.json").read())
print(sys.argv[1], "n", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "sync_bn", d["config"]["sync_bn"])
PY
}
{
  timeout 300 python bench.py --steps 12 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-config1 2>/dev/null | grep -E "^\{" > gpurun_out/r2_scale8_n1.json
  python -c "
import json; d=json.loads(open('gpurun_out/r2_scale8_n1.json').read()); print('n1 value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"
  run 8 n8
  run 8 n8_syncbn --sync-bn
  run 4 n4
} 2>&1 | tee gpurun_out/r2_scale8.txt
