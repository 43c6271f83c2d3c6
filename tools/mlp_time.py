"""Device time of every launch of one MLP forward+backward (run under ncu / gpurun)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdlrm_b200 import model_no_ddp as M  # noqa: E402

dev = torch.device("cuda:0")
B = 8192
np.random.seed(1)
net = M.DLRM_Net(np.asarray([13, 512, 256, 128]), np.asarray([479, 512, 512, 256, 1]), arch_interaction_op="dot",
                 arch_interaction_itself=False, sigmoid_bot=-1, sigmoid_top=3).to(dev)
net.mlp_impl = "tcgen05"
x = torch.randn(B, 479, device=dev)
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    xi = x.clone().requires_grad_()
    y = net.apply_mlp("top", xi)
    y.backward(torch.ones_like(y) * 0.01)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
