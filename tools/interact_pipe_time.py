"""Device time of the interaction backward at the Terabyte shape (B = 8192, 27 x 128) with padded gradient rows
(what the top MLP hands over): software-pipelined kernel (cdlrm_interact_set_option(1, 1)) vs interact_bwd_kernel.
Run under gpurun."""
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
sets = [(torch.randn(B, d, device=dev).requires_grad_(), [torch.randn(B, d, device=dev).requires_grad_() for _ in range(F - 1)])
        for _ in range(12)]
npair = F * (F - 1) // 2
dR = torch.randn(B, 480, device=dev)[:, :d + npair]
for pipe in (2, 0, 1, 2, 0):
    check(lib.cdlrm_interact_set_option(1, min(pipe, 1)))
    check(lib.cdlrm_interact_set_option(2, pipe))
    for rep in range(2):
        if rep == 1:
            lib.cdlrm_prof_enable(1)
        for x, ly in sets:
            R = net.interact_features(x, ly)
            R.backward(dR)
            lib.cdlrm_prof_null(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    ms = (C.c_double * NK)()
    calls = (C.c_int64 * NK)()
    check(lib.cdlrm_prof_report(ms, calls, NK))
    lib.cdlrm_prof_enable(0)
    out = []
    for nm in ("interact_fwd", "interact_bwd", "null"):
        i = names.index(nm)
        out.append(f"{nm}: {ms[i] * 1e3 / max(calls[i], 1):.1f} us")
    print({0: "plain              ", 1: "pipelined (2-stage)", 2: "pipelined (1-stage)"}[pipe], "  ".join(out), flush=True)
check(lib.cdlrm_interact_set_option(1, 1))
check(lib.cdlrm_interact_set_option(2, 0))
