"""Device time of the interaction kernels at the Terabyte shape (B = 8192, 27 x 128), tensor-core path vs
the CUDA-core paths (cdlrm_interact_set_option).  Run under gpurun."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdlrm_b200 import model_no_ddp as M  # noqa: E402
from cdlrm_b200._lib import check, lib  # noqa: E402

dev = torch.device("cuda:0")
B, F, d = 8192, 27, 128
net = M.DLRM_Net.__new__(M.DLRM_Net)
torch.nn.Module.__init__(net)
net.arch_interaction_op, net.arch_interaction_itself = "dot", False
NK = lib.cdlrm_prof_num_kernels()
names = [lib.cdlrm_prof_kernel_name(i).decode() for i in range(NK)]
# several input sets so that consecutive launches do not re-read the same lines from L2
sets = [(torch.randn(B, d, device=dev).requires_grad_(), [torch.randn(B, d, device=dev).requires_grad_() for _ in range(F - 1)])
        for _ in range(12)]
dR = torch.randn(B, d + F * (F - 1) // 2, device=dev)
for variant, label in ((0, "cuda-core (smem transpose, FFMA2)"), (2, "cuda-core (butterfly)"), (1, "mma.sync 3xTF32")):
    check(lib.cdlrm_interact_set_option(0, variant))
    for rep in range(2):
        if rep == 1:
            lib.cdlrm_prof_enable(1)
        for x, ly in sets:
            R = net.interact_features(x, ly)
            R.backward(dR)
    ms = (C.c_double * NK)()
    calls = (C.c_int64 * NK)()
    check(lib.cdlrm_prof_report(ms, calls, NK))
    lib.cdlrm_prof_enable(0)
    for nm in ("interact_fwd", "interact_bwd"):
        i = names.index(nm)
        print(f"{label:36s} {nm}: {ms[i] * 1e3 / max(calls[i], 1):.1f} us over {calls[i]} launches")
check(lib.cdlrm_interact_set_option(0, 0))
