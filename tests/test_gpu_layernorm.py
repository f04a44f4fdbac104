"""normalization='layer_norm' (an ablation option of the reference, gconv_adapter.py:54-61,98-106): the fused
LayerNorm + scalar pass (gca_layernorm.cu) against torch.nn.functional.layer_norm in fp64, forward and backward, and the
module with the fused tail against the module with the stock ops."""
import pytest
import torch
import torch.nn.functional as F

from gconv_adapter_b200 import GConvAdapter
from gconv_adapter_b200.finetune.gconv_adapter import _LayerNormScaleFunction
from gconv_adapter_b200.graphs.synthetic import make_inputs, symmetric_random_graph
from util import load_module_params

pytestmark = pytest.mark.gpu


def _close(a, b, tag, tol=2e-6):
    err = float((a.double() - b.double()).abs().max())
    ref = float(b.double().abs().max())
    assert err <= tol * max(ref, 1e-30) + 1e-30, f"{tag}: max err {err:.3e} vs max|ref| {ref:.3e}"


@pytest.mark.parametrize("n,d", [(1, 4), (7, 64), (1000, 300), (5000, 256), (777, 1024), (33, 516)])
@pytest.mark.parametrize("affine,with_scalar", [(True, True), (True, False), (False, True)])
def test_fused_layernorm_scale_matches_fp64(n, d, affine, with_scalar):
    g = torch.Generator().manual_seed(n * 31 + d)
    x = (torch.randn(n, d, generator=g) * 3 + 0.5).cuda()
    gy = torch.randn(n, d, generator=g).cuda()
    w = (1 + 0.3 * torch.randn(d, generator=g)).cuda() if affine else None
    b = (0.2 * torch.randn(d, generator=g)).cuda() if affine else None
    s = torch.tensor([1.3]).cuda() if with_scalar else None
    leaves = [t.clone().requires_grad_(True) if t is not None else None for t in (x, w, b, s)]
    y = _LayerNormScaleFunction.apply(leaves[0], leaves[1], leaves[2], leaves[3], 1e-5)
    y.backward(gy)
    ref = [t.double().clone().requires_grad_(True) if t is not None else None for t in (x, w, b, s)]
    yr = F.layer_norm(ref[0], (d,), ref[1], ref[2], 1e-5)
    if ref[3] is not None:
        yr = yr * ref[3]
    yr.backward(gy.double())
    _close(y, yr.detach(), "y")
    _close(leaves[0].grad, ref[0].grad, "g_x", tol=5e-6)
    for k, name in ((1, "g_weight"), (2, "g_bias")):
        if leaves[k] is not None:
            _close(leaves[k].grad, ref[k].grad, name, tol=5e-6)
    if s is not None:
        floor = 5e-7 * float((gy.abs().double() * yr.detach().abs()).sum()) / 1.3
        assert abs(float(leaves[3].grad) - float(ref[3].grad)) <= 1e-5 * abs(float(ref[3].grad)) + floor
    # fixed-order reductions: a second run gives the same bits
    again = [t.clone().requires_grad_(True) if t is not None else None for t in (x, w, b, s)]
    y2 = _LayerNormScaleFunction.apply(again[0], again[1], again[2], again[3], 1e-5)
    y2.backward(gy)
    assert torch.equal(y2, y) and torch.equal(again[0].grad, leaves[0].grad)
    for k in (1, 2, 3):
        if leaves[k] is not None:
            assert torch.equal(again[k].grad, leaves[k].grad)


@pytest.mark.parametrize("three_d", [False, True])
@pytest.mark.parametrize("n,d,r", [(3000, 64, 16), (20000, 256, 16)])
def test_module_with_fused_tail_equals_stock_ops(n, d, r, three_d):
    ei = symmetric_random_graph(n, 6 * n, seed=3).cuda()
    x, g_out, params = make_inputs(n, d, r, seed=4)
    outs = []
    for fuse in (True, False):
        m = GConvAdapter(d, r, normalization="layer_norm", learnable_scalar=True)
        load_module_params(m, params)
        with torch.no_grad():
            m.normalization.weight.copy_(1 + 0.1 * torch.arange(d) / d)
            m.normalization.bias.copy_(0.05 * torch.ones(d))
        m = m.cuda()
        m.fuse_layer_norm = fuse
        xx = x.cuda().requires_grad_(True)
        inp = xx.unsqueeze(0) if three_d else xx
        y = m(inp, ei)
        assert y.shape == inp.shape
        y.backward(g_out.cuda().unsqueeze(0) if three_d else g_out.cuda())
        outs.append([y.detach(), xx.grad] + [p.grad for _, p in sorted(m.named_parameters())])
    names = ["y", "g_x"] + [k for k, _ in sorted(GConvAdapter(d, r, normalization="layer_norm").named_parameters())]
    for k, a, b in zip(names, outs[0], outs[1]):
        tol = 2e-5 if k == "scalar" else 5e-6
        _close(a, b, k, tol=tol)
