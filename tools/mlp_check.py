"""Quick GPU check + timing of the tensor-core MLP path against torch (run under gpurun)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdlrm_b200 import model_no_ddp as M  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
np.random.seed(1)
net = M.DLRM_Net(np.asarray([13, 512, 256, 128]), np.asarray([479, 512, 512, 256, 1]), arch_interaction_op="dot",
                 arch_interaction_itself=False, sigmoid_bot=-1, sigmoid_top=3).to(dev)
for which, K in (("bot", 13), ("top", 479)):
    seq = net.bot_l if which == "bot" else net.top_l
    x = torch.randn(B, K, device=dev)
    res = {}
    for impl in ("torch", "tcgen05"):
        net.mlp_impl = impl
        xi = x.clone().requires_grad_()
        for p in seq.parameters():
            p.grad = None
        y = net.apply_mlp(which, xi)
        dy = torch.ones_like(y) * 0.01
        y.backward(dy)
        torch.cuda.synchronize()
        res[impl] = (y.detach(), xi.grad, [p.grad.clone() for p in seq.parameters()])
        # timing
        for _ in range(3):
            xi = x.clone().requires_grad_()
            y = net.apply_mlp(which, xi)
            y.backward(dy)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            xi = x.clone().requires_grad_()
            y = net.apply_mlp(which, xi)
            y.backward(dy)
        e1.record()
        torch.cuda.synchronize()
        print(f"{which} {impl}: {e0.elapsed_time(e1) / 20 * 1000:.1f} us per fwd+bwd (eager, B={B})")
    a, b = res["tcgen05"], res["torch"]

    def rel(u, v):
        return float((u - v).abs().max() / v.abs().max().clamp_min(1e-30))
    print(which, "y", rel(a[0], b[0]), "dx", rel(a[1], b[1]), "grads", [f"{rel(u, v):.2e}" for u, v in zip(a[2], b[2])])
