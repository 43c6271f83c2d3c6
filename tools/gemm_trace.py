"""Timeline of CTA 0 of one forward GEMM (cdlrm_mlp_set_trace): when each k-block's TMA was issued, when its
operands had landed, when each K segment's accumulator was ready, when the last store was done.  Run under gpurun."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdlrm_b200 import model_no_ddp as M  # noqa: E402
from cdlrm_b200._lib import check, lib  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
Mrows, K, N = 8192, 512, 512
lin = torch.nn.Linear(K, N).to(dev)
st = M._MlpState(torch.nn.Sequential(lin), -2)
x = torch.randn(Mrows, K, device=dev)
buf = torch.zeros(4, 128, dtype=torch.int64, device=dev)
for dbg in (0, 4):
    check(lib.cdlrm_mlp_set_option(3, dbg))
    for _ in range(3):
        M._MlpFn.apply(st, x, lin.weight, lin.bias)
    torch.cuda.synchronize()
    buf.zero_()
    check(lib.cdlrm_mlp_set_trace(buf.data_ptr()))
    M._MlpFn.apply(st, x, lin.weight, lin.bias)
    torch.cuda.synchronize()
    check(lib.cdlrm_mlp_set_trace(None))
    t = buf.cpu()
    t0 = int(t[3, 0])
    rel = lambda r: [int(v) - t0 for v in t[r] if int(v) > 0]
    print(f"dbg {dbg}: times in ns after the prologue")
    print("  tma issue :", rel(0)[:72])
    print("  landed    :", rel(1)[:72])
    print("  acc ready :", rel(2)[:8])
    print("  done      :", int(t[3, 1]) - t0)
check(lib.cdlrm_mlp_set_option(3, 0))
