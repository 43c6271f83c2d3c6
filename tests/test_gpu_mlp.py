"""GPU parity of the tensor-core MLP path (cdlrm_mlp_*: 3xTF32 split GEMMs, csrc/mlp.cu) against
the stock nn.Sequential the reference builds in DLRM_Net.create_mlp (model_no_ddp.py:244-270).

Floating point, so the bar is the north_star tolerance (1e-5 relative to the tensor's scale,
util.assert_close_fp32) against an FP64 evaluation of the same layers, and -- to show that the
split products really deliver FP32 accuracy -- the error must stay within a small factor of
the error torch's own FP32 (SIMT sgemm) path makes against the same FP64 reference."""
import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _net(ln_bot, ln_top, impl):
    from cdlrm_b200 import model_no_ddp as M
    np.random.seed(17)
    net = M.DLRM_Net(np.asarray(ln_bot), np.asarray(ln_top), arch_interaction_op="dot", arch_interaction_itself=False,
                     sigmoid_bot=-1, sigmoid_top=len(ln_top) - 2).to(DEV)
    net.mlp_impl = impl
    return net


def _seq64(seq):
    seq64 = torch.nn.Sequential(*[type(m)(m.in_features, m.out_features) if isinstance(m, torch.nn.Linear) else type(m)()
                                  for m in seq]).double().to(DEV)
    seq64.load_state_dict({k: v.double() for k, v in seq.state_dict().items()})
    return seq64


def _ambiguous_rows(seq, x, thr=1e-4):
    """Rows with a ReLU pre-activation within `thr` of zero somewhere in the stack: there the ReLU
    mask (hence the whole gradient of the row) legitimately depends on the last bits of the forward,
    for ANY FP32 implementation.  The caller zeroes dy on these rows so that gradients compare."""
    amb = torch.zeros(x.shape[0], dtype=torch.bool, device=x.device)
    h = x.double()
    mods = list(_seq64(seq))
    for i, m in enumerate(mods):
        h = m(h)
        if isinstance(m, torch.nn.Linear) and i + 1 < len(mods) and isinstance(mods[i + 1], torch.nn.ReLU):
            amb |= (h.abs() < thr).any(dim=1)
    return amb


def _ref64(seq, x, dy):
    seq64 = _seq64(seq)
    x64 = x.double().requires_grad_()
    y = seq64(x64)
    y.backward(dy.double())
    return y.detach(), x64.grad, [p.grad for p in seq64.parameters()]


def _err(a, ref):
    ref = ref.double()
    return float((a.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("which,dims,B", [
    ("bot", [13, 512, 256, 128], 8192),
    ("top", [479, 512, 512, 256, 1], 8192),
    ("top", [479, 512, 512, 256, 1], 1000),       # ragged batch: partial M tile
    ("bot", [13, 64, 16], 77),                     # configs[0]-sized layers, N below one tile
    ("top", [367, 512, 256, 1], 2048),             # configs[1] (Kaggle shape) top MLP
])
def test_mlp_matches_fp64_reference(which, dims, B):
    _check_mlp(which, dims, B)


def _check_mlp(which, dims, B):
    ln_bot = dims if which == "bot" else [13, 32, 16]
    ln_top = dims if which == "top" else [40, 8, 1]
    net = _net(ln_bot, ln_top, "tcgen05")
    seq = net.bot_l if which == "bot" else net.top_l
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(B, dims[0], device=DEV, generator=g)
    if which == "top":
        x = x * 3.0
    dy = torch.randn(B, dims[-1], device=DEV, generator=g)
    amb = _ambiguous_rows(seq, x)
    assert float(amb.float().mean()) < 0.5
    dy[amb] = 0.0
    y64, dx64, gp64 = _ref64(seq, x, dy)

    def run(impl):
        net.mlp_impl = impl
        for p in seq.parameters():
            p.grad = None
        xi = x.clone().requires_grad_()
        y = net.apply_mlp(which, xi)
        y.backward(dy)
        torch.cuda.synchronize()
        return y.detach(), xi.grad, [p.grad.clone() for p in seq.parameters()]

    y_t, dx_t, gp_t = run("torch")
    y_c, dx_c, gp_c = run("tcgen05")
    pairs = [("y", y_c, y_t, y64), ("dx", dx_c, dx_t, dx64)] + \
            [(f"grad{i}", a, b, c) for i, (a, b, c) in enumerate(zip(gp_c, gp_t, gp64))]
    for name, mine, theirs, ref in pairs:
        e_mine, e_torch = _err(mine, ref), _err(theirs, ref)
        assert e_mine <= 1e-5, f"{which} {name}: {e_mine:.3e} off the FP64 reference (torch FP32: {e_torch:.3e})"
        # bias gradients of narrow layers are sums of B signed terms that cancel (a single scalar for the
        # last top layer): the error is relative to sum |term|, about sqrt(B) times the result, so the
        # floor is 1e-6 of the result's scale there, 2e-7 elsewhere
        floor = 1e-6 if mine.numel() <= 16 else 2e-7
        assert e_mine <= 8 * e_torch + floor, f"{which} {name}: {e_mine:.3e} vs torch FP32 {e_torch:.3e}: not FP32-grade"
        util.assert_close_fp32(mine.cpu().numpy(), ref.float().cpu().numpy(), err_msg=f"{which} {name}")


def test_dlrm_step_same_loss_and_grads_both_mlp_paths():
    """One full DLRM_Net forward/backward (bottom MLP -> interaction -> top MLP -> BCE) with the
    tensor-core MLPs against the stock Sequential path: loss and every parameter gradient."""
    ln_bot, ln_top = [13, 512, 256, 128], [479, 512, 512, 256, 1]
    B, d, T = 2048, 128, 26
    net = _net(ln_bot, ln_top, "torch")
    g = torch.Generator(device=DEV).manual_seed(1)
    X = torch.randn(B, 13, device=DEV, generator=g)
    ly = [torch.randn(B, d, device=DEV, generator=g) * 0.05 for _ in range(T)]
    Y = (torch.rand(B, 1, device=DEV, generator=g) < 0.25).float()
    out = {}
    for impl in ("torch", "tcgen05"):
        net.mlp_impl = impl
        net.zero_grad(set_to_none=True)
        lyi = [t.clone().requires_grad_() for t in ly]
        loss = torch.nn.functional.binary_cross_entropy(net(X, lyi), Y)
        loss.backward()
        out[impl] = (loss.item(), [p.grad.clone() for p in net.parameters()], [t.grad.clone() for t in lyi])
    assert abs(out["tcgen05"][0] - out["torch"][0]) <= 1e-5 * abs(out["torch"][0])
    # gradients: a ReLU unit whose pre-activation is within rounding of zero may fire in one path and
    # not in the other (a few of 10 M units per step); that moves single rows, not the bulk
    # parameter gradients are sums over the batch: the handful of flipped units moves them by O(flips / B)
    names = [n for n, _ in net.named_parameters()] + [f"ly{k}" for k in range(T)]
    for nm, a, b in zip(names, out["tcgen05"][1] + out["tcgen05"][2], out["torch"][1] + out["torch"][2]):
        rel = float((a - b).norm() / b.norm().clamp_min(1e-30))
        assert rel <= 5e-3, f"{nm}: relative Frobenius difference {rel:.3e}"
        if nm.startswith("ly"):      # per-sample gradients: all but the flipped rows agree to FP32 accuracy
            scale = float(b.abs().max())
            bad_rows = ((a - b).abs().amax(dim=1) > 2e-5 * scale).float().mean().item()
            assert bad_rows <= 5e-3, f"{nm}: {bad_rows:.2%} of the rows off by more than 2e-5 of the scale"


def test_flat_bucket_mode_same_gradients_and_sgd_step():
    """DLRM_Net.flatten_parameters (one flat parameter buffer + one gradient bucket, one-launch SGD) against the
    per-tensor autograd path + torch.optim.SGD (main_no_ddp.py:375,413) on the same inputs."""
    ln_bot, ln_top = [13, 512, 256, 128], [479, 512, 512, 256, 1]
    B, d, T, lr = 1024, 128, 26, 0.1
    g = torch.Generator(device=DEV).manual_seed(3)
    X = torch.randn(B, 13, device=DEV, generator=g)
    ly = [torch.randn(B, d, device=DEV, generator=g) * 0.05 for _ in range(T)]
    Y = (torch.rand(B, 1, device=DEV, generator=g) < 0.25).float()
    ref, flat = _net(ln_bot, ln_top, "tcgen05"), _net(ln_bot, ln_top, "tcgen05")
    flat.flatten_parameters()
    opt = torch.optim.SGD(ref.parameters(), lr=lr)
    for step in range(2):
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.binary_cross_entropy(ref(X, ly), Y).backward()
        torch.nn.functional.binary_cross_entropy(flat(X, ly), Y).backward()
        for (n, a), b in zip(ref.named_parameters(), flat.parameters()):
            assert b.grad.data_ptr() >= flat.flat_grads.data_ptr()
            util.assert_close_fp32(b.grad.cpu().numpy(), a.grad.cpu().numpy(), err_msg=f"step {step} grad {n}")
        opt.step()
        flat.flat_sgd_step(lr)
        for (n, a), b in zip(ref.named_parameters(), flat.parameters()):
            util.assert_close_fp32(b.detach().cpu().numpy(), a.detach().cpu().numpy(), err_msg=f"step {step} param {n}")


@pytest.mark.parametrize("B", [1, 77, 8192, 65536])
def test_fused_bce_mean_matches_torch(B):
    """cdlrm_bce_mean (loss + derivative in one launch) against torch.nn.BCELoss(reduction="mean")
    (main_no_ddp.py:355-369), including saturated probabilities (log clamp at -100, derivative floor)."""
    from cdlrm_b200 import model_no_ddp as M
    g = torch.Generator(device=DEV).manual_seed(B)
    z = torch.sigmoid(torch.randn(B, 1, device=DEV, generator=g) * 4)
    if B > 4:
        z[0], z[1], z[2] = 0.0, 1.0, 1e-30
    t = (torch.rand(B, 1, device=DEV, generator=g) < 0.25).float()
    z1, z2 = z.clone().requires_grad_(), z.clone().requires_grad_()
    l1 = torch.nn.BCELoss(reduction="mean")(z1, t)
    l2 = M.bce_mean(z2, t)
    (l1 * 3.0).backward()
    (l2 * 3.0).backward()
    assert abs(l1.item() - l2.item()) <= 1e-5 * abs(l1.item())
    util.assert_close_fp32(z2.grad.cpu().numpy(), z1.grad.cpu().numpy())


def test_mlp_cluster_multicast_mode_matches_fp64_reference():
    """The optional CTA-pair mode of the GEMM (cdlrm_mlp_set_option(4, 2): the two CTAs of a cluster share every
    B tile through TMA multicast, stage release through a multicast tcgen05.commit) must give the same results."""
    from cdlrm_b200._lib import check, lib
    check(lib.cdlrm_mlp_set_option(4, 2))
    try:
        _check_mlp("top", [479, 512, 512, 256, 1], 8192)
        _check_mlp("bot", [13, 512, 256, 128], 1000)       # 8 m tiles: pairs; ragged last tile
    finally:
        check(lib.cdlrm_mlp_set_option(4, 1))
