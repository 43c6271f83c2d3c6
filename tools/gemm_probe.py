"""Where the time of one 3xTF32 GEMM goes: device time of the forward GEMM of one Linear layer with parts of
the kernel switched off (cdlrm_mlp_set_option(3, bits): 1 no result stores, 2 no lo-operand loads, 4 no MMAs;
results are wrong with any bit set -- measurement only).  Run under gpurun."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdlrm_b200 import model_no_ddp as M  # noqa: E402
from cdlrm_b200._lib import check, lib  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
NK = lib.cdlrm_prof_num_kernels()
names = [lib.cdlrm_prof_kernel_name(i).decode() for i in range(NK)]


def gemm_us(Mrows, K, N, reps=20):
    lin = torch.nn.Linear(K, N).to(dev)
    st = M._MlpState(torch.nn.Sequential(lin), -2)
    x = torch.randn(Mrows, K, device=dev)
    for _ in range(3):
        M._MlpFn.apply(st, x, lin.weight, lin.bias)
    torch.cuda.synchronize()
    lib.cdlrm_prof_enable(1)
    for _ in range(reps):
        M._MlpFn.apply(st, x, lin.weight, lin.bias)
    ms = (C.c_double * NK)()
    calls = (C.c_int64 * NK)()
    check(lib.cdlrm_prof_report(ms, calls, NK))
    lib.cdlrm_prof_enable(0)
    i = names.index("mlp_gemm")
    return ms[i] * 1e3 / max(calls[i], 1)


for shape in ((8192, 512, 512), (8192, 512, 256), (8192, 256, 128)):
    row = []
    for dbg in (0, 1, 2, 3, 4, 5, 6, 7):
        check(lib.cdlrm_mlp_set_option(3, dbg))
        row.append(f"{dbg}:{gemm_us(*shape):6.1f}")
    print(shape, "  ".join(row), flush=True)
check(lib.cdlrm_mlp_set_option(3, 0))
