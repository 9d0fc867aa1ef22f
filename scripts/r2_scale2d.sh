#!/bin/bash
# N = 2 sanity of the final round-2 state (weight bank + fused SGD + prefetching e2e loop under DDP)
mkdir -p gpurun_out
run() {
  label=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 12 --warmup 5 "$@" 2>gpurun_out/r2_scale2d_$label.err | grep -E "^\{" > gpurun_out/r2_scale2d_$label.json
  python - "$label" <<'PY' || tail -5 gpurun_out/r2_scale2d_$label.err
import json, sys
d = json.loads(open(f"gpurun_out/r2_scale2d_{sys.argv[1]}.json").read())
print(sys.argv[1], "n", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "sync_bn", d["config"].get("sync_bn"))
PY
}
{
  timeout 300 python bench.py --steps 12 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-config1 2>/dev/null | grep -E "^\{" > gpurun_out/r2_scale2d_n1.json
  python -c "
import json; d=json.loads(open('gpurun_out/r2_scale2d_n1.json').read()); print('n1 value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"
  run n2
  run n2_syncbn --sync-bn
} 2>&1 | tee gpurun_out/r2_scale2d.txt
