"""Device time of the dim-128 interaction forward kernels at the Terabyte shape (B = 8192, 27 x 128):
0 = interact_fwd_tr_kernel (a warp per sample), 3 = interact_fwd_h_kernel (half a warp per sample), 1 / 2 = the
software-pipelined ring variants, 5 / 6 = interact_fwd_hs_kernel (half-warp arithmetic fed through shared memory).  CUDA events around every launch (cdlrm_prof_*), 12 input sets (> L2).  Run under gpurun."""
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
sets = [(torch.randn(B, d, device=dev), [torch.randn(B, d, device=dev) for _ in range(F - 1)]) for _ in range(12)]
s = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for fwd, stag in ((0, 0), (3, 0), (5, 0), (6, 0), (4, 800), (3, 0), (5, 0), (6, 0)):
    check(lib.cdlrm_interact_set_option(2, fwd))
    check(lib.cdlrm_interact_set_option(3, stag))
    for rep in range(3):
        if rep == 1:
            lib.cdlrm_prof_enable(1)
        torch.cuda._sleep(20_000_000)           # the host enqueues everything while the GPU is parked
        with torch.no_grad():
            for x, ly in sets:
                net.interact_features(x, ly)
                lib.cdlrm_prof_null(s)
    ms = (C.c_double * NK)()
    calls = (C.c_int64 * NK)()
    check(lib.cdlrm_prof_report(ms, calls, NK))
    lib.cdlrm_prof_enable(0)
    i, j = names.index("interact_fwd"), names.index("null")
    raw = ms[i] * 1e3 / max(calls[i], 1)
    null = ms[j] * 1e3 / max(calls[j], 1)
    algo = B * (F * 4 * d + (d + F * (F - 1) // 2) * 4)
    print(f"fwd variant {fwd} stagger {stag} ns: {raw:.1f} us raw, {raw - null:.1f} us net of the event pair ({null:.1f}); "
          f"{algo / (raw - null) / 1e3:.0f} GB/s algorithmic over {calls[i]} launches")
check(lib.cdlrm_interact_set_option(2, -1))
