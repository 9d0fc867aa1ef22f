"""TEST INFRASTRUCTURE ONLY: deterministic parameter values derived from the state_dict KEY, so that the
reference model (when the golden is generated) and any model under test can be given identical weights
without storing 150 MB of them."""
import zlib

import torch


def fill_(state: dict) -> dict:
    """Returns {key: tensor} with the shapes / dtypes of ``state`` and key-seeded values."""
    out = {}
    for key, ref in state.items():
        g = torch.Generator().manual_seed(zlib.crc32(key.encode()))
        if key.endswith("num_batches_tracked"):
            val = torch.zeros_like(ref)
        elif key.endswith("running_mean"):
            val = torch.zeros_like(ref)
        elif key.endswith("running_var"):
            val = torch.ones_like(ref)
        elif ref.dim() >= 2:                                   # conv kernels [K, Cin, Cout] / [Cin, Cout], linear
            fan_in = ref[..., 0].numel() if key.endswith("kernel") else ref.shape[1]
            val = torch.randn(ref.shape, generator=g) / fan_in ** 0.5
        elif key.endswith("weight"):                           # batch-norm scale
            val = 1.0 + 0.1 * torch.randn(ref.shape, generator=g)
        else:                                                  # biases
            val = 0.1 * torch.randn(ref.shape, generator=g)
        out[key] = val.to(ref.dtype)
    return out
