#!/usr/bin/env python
"""CIN accuracy probe: forward and backward of the tensor-core path against a float64 einsum on the GPU, at config-3 layer
shapes with a reduced batch.  Prints max-norm relative error and the element-wise excess (|a-b| / (1e-5|ref| + 1e-5 rms))."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from recalgorithm_b200 import ops  # noqa: E402


def err(a, b):
    a, b = a.double(), b.double()
    rms = b.pow(2).mean().sqrt()
    return {"maxnorm": float((a - b).abs().max() / b.abs().max()), "elementwise_excess": float(((a - b).abs() / (1e-5 * b.abs() + 1e-5 * rms)).max())}


def main():
    torch.manual_seed(0)
    for (B, m, hk, D, H) in ((256, 30, 30, 16, 128), (256, 30, 128, 16, 128)):
        x0 = torch.randn((B, m, D), device="cuda") * 0.25
        xk = torch.randn((B, hk, D), device="cuda") * 0.25
        w = torch.randn((hk * m, H), device="cuda") * 0.05
        g = torch.randn((B, H, D), device="cuda")
        x0d, xkd, wd, gd = (t.double().requires_grad_() for t in (x0, xk, w, g))
        ref = torch.einsum("bid,bjd,ijn->bnd", xkd, x0d, wd.reshape(hk, m, H))
        (ref * gd.detach()).sum().backward()
        out = ops.cin_fwd(x0, xk, w)
        dx0, dxk, dw = ops.cin_bwd(x0, xk, w, g)
        print(json.dumps({"shape": [B, m, hk, D, H], "fwd": err(out, ref.detach()), "dx0": err(dx0, x0d.grad), "dxk": err(dxk, xkd.grad),
                          "dw": err(dw, wd.grad)}), flush=True)


if __name__ == "__main__":
    main()
