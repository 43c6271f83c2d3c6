"""Error of the 3xTF32 tensor-core GEMM against FP64, next to torch's FP32 matmul, for the knobs of
cdlrm_mlp_set_option (split rounding, K segment per TMEM accumulation chain).  Run under gpurun."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdlrm_b200 import model_no_ddp as M  # noqa: E402
from cdlrm_b200._lib import check, lib  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def one(K, N, Mrows, trunc, seg):
    check(lib.cdlrm_mlp_set_option(0, trunc))
    check(lib.cdlrm_mlp_set_option(1, seg))
    lin = torch.nn.Linear(K, N).to(dev)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(N, K, device=dev) / K ** 0.5)
        lin.bias.zero_()
    seq = torch.nn.Sequential(lin)
    st = M._MlpState(seq, -2)          # -2: no activation on the last layer
    x = torch.randn(Mrows, K, device=dev)
    y = M._MlpFn.apply(st, x, lin.weight, lin.bias)
    ref = x.double() @ lin.weight.double().t()
    yt = x @ lin.weight.t()
    torch.cuda.synchronize()
    sc = ref.abs().max()
    return (float((y.double() - ref).abs().max() / sc), float((y.double() - ref).pow(2).mean().sqrt() / sc),
            float((yt.double() - ref).abs().max() / sc), float((yt.double() - ref).pow(2).mean().sqrt() / sc))


print("K      trunc  seg   max_err     rms_err    | torch max   torch rms")
for K in (32, 128, 512, 2048, 8192):
    for trunc, seg in ((1, 0), (0, 0), (0, 16), (0, 8), (0, 4)):
        e = one(K, 256, 1024, trunc, seg)
        print(f"{K:6d} {trunc:5d} {seg:4d}   {e[0]:.3e}  {e[1]:.3e}  | {e[2]:.3e}  {e[3]:.3e}")
check(lib.cdlrm_mlp_set_option(0, 0))
check(lib.cdlrm_mlp_set_option(1, 8))
