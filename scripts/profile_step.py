"""Host/device breakdown of one MinkUNet-34 AMP training step (torch.profiler)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import openpcseg_b200.torchsparse as ts  # noqa: E402
from openpcseg_b200.segmentors import MinkUNet, minkunet34_config  # noqa: E402
from openpcseg_b200.synthetic import make_batch  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda")
torch.manual_seed(0)
model = MinkUNet(minkunet34_config()).to(dev).train()
opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, nesterov=True)
scaler = torch.amp.GradScaler("cuda")
b = make_batch(list(range(batch)))
data = {k: torch.from_numpy(b[k]).to(dev) for k in ("coords", "feats", "labels")}


def phase_step(timed=False):
    t = {}
    def mark(name, t0):
        if timed:
            torch.cuda.synchronize()
            t[name] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    lidar = ts.SparseTensor(data["feats"], data["coords"], 1)
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.float16):
        logits = model.forward_logits(lidar)
    mark("forward_logits", t0)
    t0 = time.perf_counter()
    loss = model.criterion(logits, data["labels"])
    mark("loss", t0)
    t0 = time.perf_counter()
    scaler.scale(loss).backward()
    mark("backward", t0)
    t0 = time.perf_counter()
    scaler.unscale_(opt)
    torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
    scaler.step(opt)
    scaler.update()
    mark("optimizer", t0)
    return t


for _ in range(3):
    phase_step()
torch.cuda.synchronize()
for _ in range(2):
    print("phases(ms, synced):", {k: round(v, 2) for k, v in phase_step(True).items()})
# un-synced host time of a step (how long the CPU needs to enqueue everything)
torch.cuda.synchronize()
t0 = time.perf_counter()
phase_step()
host = (time.perf_counter() - t0) * 1e3
torch.cuda.synchronize()
total = (time.perf_counter() - t0) * 1e3
print(f"host enqueue {host:.1f} ms, step total {total:.1f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        phase_step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=28, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=22, max_name_column_width=60))
