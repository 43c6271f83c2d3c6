"""CTA-pair multicast of the B tiles (cdlrm_mlp_set_option(4, 2)) against single CTAs (4, 1): device time of the
forward GEMM of single layers and of all GEMMs of one top-MLP forward + backward.  Run under gpurun."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdlrm_b200 import model_no_ddp as M  # noqa: E402
from cdlrm_b200._lib import check, lib  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
NK = lib.cdlrm_prof_num_kernels()
names = [lib.cdlrm_prof_kernel_name(i).decode() for i in range(NK)]


def report():
    ms = (C.c_double * NK)()
    calls = (C.c_int64 * NK)()
    check(lib.cdlrm_prof_report(ms, calls, NK))
    i = names.index("mlp_gemm")
    return ms[i] * 1e3, calls[i]


def layer_us(Mrows, K, N, reps=20):
    lin = torch.nn.Linear(K, N).to(dev)
    st = M._MlpState(torch.nn.Sequential(lin), -2)
    x = torch.randn(Mrows, K, device=dev)
    for _ in range(3):
        M._MlpFn.apply(st, x, lin.weight, lin.bias)
    torch.cuda.synchronize()
    lib.cdlrm_prof_enable(1)
    for _ in range(reps):
        M._MlpFn.apply(st, x, lin.weight, lin.bias)
    us, n = report()
    lib.cdlrm_prof_enable(0)
    return us / max(n, 1)


def mlp_us(which, reps=10):
    np.random.seed(1)
    net = M.DLRM_Net(np.asarray([13, 512, 256, 128]), np.asarray([479, 512, 512, 256, 1]), arch_interaction_op="dot",
                     arch_interaction_itself=False, sigmoid_bot=-1, sigmoid_top=3).to(dev)
    x = torch.randn(8192, 479 if which == "top" else 13, device=dev)
    for it in range(3 + reps):
        if it == 3:
            torch.cuda.synchronize()
            lib.cdlrm_prof_enable(1)
        xi = x.clone().requires_grad_()
        y = net.apply_mlp(which, xi)
        y.backward(torch.ones_like(y) * 0.01)
    us, n = report()
    lib.cdlrm_prof_enable(0)
    return us / reps, n / reps


for cl in (1, 2):
    check(lib.cdlrm_mlp_set_option(4, cl))
    row = [f"{s}: {layer_us(*s):6.1f}" for s in ((8192, 512, 512), (8192, 512, 256), (8192, 256, 128), (8192, 479, 512))]
    t_top, n_top = mlp_us("top")
    t_bot, n_bot = mlp_us("bot")
    print(f"cluster {cl}:", "  ".join(row), f" | top MLP fwd+bwd {t_top:7.1f} us in {n_top:.0f} GEMMs, bottom {t_bot:7.1f} us in {n_bot:.0f}",
          flush=True)
