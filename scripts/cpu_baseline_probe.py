"""Time the reference CPU path (oracle/_ref) on a small sub-scan for several thread counts.
Usage: python scripts/cpu_baseline_probe.py <threads> <n_azimuth>"""
import os
import sys
import time

threads, n_az = int(sys.argv[1]), int(sys.argv[2])
if threads > 0:
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ["MKL_NUM_THREADS"] = str(threads)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

if threads > 0:
    torch.set_num_threads(threads)
from oracle.cpu_minkunet import CpuMinkUNet, kind  # noqa: E402
from openpcseg_b200.segmentors import MinkUNet, minkunet34_config  # noqa: E402
from openpcseg_b200.synthetic import make_batch  # noqa: E402

torch.manual_seed(0)
state = MinkUNet(minkunet34_config()).state_dict()
b = make_batch([0], n_azimuth=n_az)
net = CpuMinkUNet(state)
t0 = time.perf_counter()
_, loss = net.forward(torch.from_numpy(b["coords"]), torch.from_numpy(b["feats"]), torch.from_numpy(b["labels"]))
t1 = time.perf_counter()
loss.backward()
t2 = time.perf_counter()
print(f"kind={kind()} cpu_count={os.cpu_count()} threads={torch.get_num_threads()} n_az={n_az} "
      f"voxels={b['coords'].shape[0]} fwd={t1 - t0:.2f}s bwd={t2 - t1:.2f}s")
