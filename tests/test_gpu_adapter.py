"""GPU parity of the adapter forward + backward, called through the drop-in nn.Module (which goes
through the C ABI), against (a) the committed golden vectors made with the reference's own glue
class and (b) the CPU oracle on seeded inputs shaped like BASELINE.json's configs."""
import numpy as np
import pytest
import torch

import gconv_adapter_b200.finetune.gconv_adapter as impl
from gconv_adapter_b200 import GConvAdapter, GraphCache
from gconv_adapter_b200.graphs.synthetic import make_graph, make_inputs, molecule_batch, symmetric_random_graph
from oracle.pyg_restated import GConvAdapterRef

from util import PARAM_KEYS, assert_close, ctor_kwargs, golden_names, load_golden, load_module_params

pytestmark = pytest.mark.gpu


def run_ours(x, ei, params, g_out, three_d=False, **kw):
    n, d = x.shape
    r = params["conv_down.lin.weight"].shape[0]
    m = GConvAdapter(d, r, **kw)
    load_module_params(m, params)
    m = m.cuda().train()
    m.graph_cache = GraphCache()
    xin = x.cuda()
    if three_d:
        xin = xin.unsqueeze(0)
    xin = xin.clone().requires_grad_(True)
    x_before = xin.detach().clone()
    impl.DEBUG_KEEP_SAVED = True
    y = m(xin, ei.cuda())
    impl.DEBUG_KEEP_SAVED = False
    assert y.shape == xin.shape and y.data_ptr() != xin.data_ptr()
    y.backward(g_out.cuda().reshape(y.shape))
    assert torch.equal(xin.detach(), x_before), "the adapter must not modify its input"
    grads = {"x": xin.grad.reshape(n, d)}
    for k, p in m.named_parameters():
        grads[k] = p.grad
    return y.detach().reshape(n, d), grads, m


def run_oracle(x, ei, params, g_out, mask=None, **kw):
    n, d = x.shape
    r = params["conv_down.lin.weight"].shape[0]
    m = GConvAdapterRef(d, r, **kw)
    load_module_params(m, params)
    m.train()
    m.relu_mask_override = mask
    xx = x.clone().requires_grad_(True)
    y = m(xx, ei)
    y.backward(g_out)
    grads = {"x": xx.grad}
    for k, p in m.named_parameters():
        grads[k] = p.grad
    return y.detach(), grads, m


def compare(ours, ref, tag, g_out=None, batchnorm=False):
    y, g, _ = ours
    yr, gr, mref = ref
    assert_close(y, yr, f"{tag}: y")
    for k in gr:
        floor = 1e-6 * float(g_out.abs().sum(0).max()) if (batchnorm and k == "conv_up.bias") else 0.0
        if k == "scalar" and g_out is not None and getattr(mref, "normalization", None) is None:
            # d loss / d scalar = <g_out, y / s>: n*d products of either sign that cancel to ~sqrt(n*d) of their absolute
            # sum, so BOTH fp32 evaluations (ours and the oracle's) carry summation noise proportional to
            # sum |g_out * y / s|, not to the result (5e-7 ~ a few fp32 ulps of that sum)
            floor = 5e-7 * float((g_out.abs().double() * yr.abs().double()).sum()) / max(abs(float(mref.scalar)), 1e-30)
        assert_close(g[k], gr[k], f"{tag}: grad {k}", noise_floor=floor)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", [n for n in golden_names() if n != "cora_shaped"])
def test_golden_vectors(name):
    g = load_golden(name)
    cfg = g["cfg"]
    params = {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")}
    y, grads, _ = run_ours(torch.from_numpy(g["x"]), torch.from_numpy(g["edge_index"]), params,
                           torch.from_numpy(g["g_out"]), three_d=bool(cfg.get("three_d", False)), **ctor_kwargs(cfg))
    assert_close(y, g["y"], name + ": y")
    assert_close(grads["x"], g["g_x"], name + ": g_x")
    for k in params:
        # BatchNorm removes the column mean, so d loss / d conv_up.bias is exactly 0 in real arithmetic:
        # both sides hold only cancellation noise of sum_i |g_out[i]| * eps.
        floor = 1e-6 * float(np.abs(g["g_out"]).sum(0).max()) if (name == "mid_batchnorm" and k == "conv_up.bias") else 0.0
        assert_close(grads[k], g["grad." + k], f"{name}: grad {k}", noise_floor=floor)


def test_golden_cora_shaped_config0():
    """BASELINE.json configs[0]: 2,708 nodes, 10,556 edges, hidden 64, rank 8."""
    g = load_golden("cora_shaped")
    ei = symmetric_random_graph(2708, 10556, seed=int(g["graph_seed"]))
    x, g_out, _ = make_inputs(2708, 64, 8, seed=int(g["input_seed"]))
    params = {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")}
    y, grads, _ = run_ours(x, ei, params, g_out, **ctor_kwargs(g["cfg"]))
    rows = torch.from_numpy(g["rows"])
    assert_close(y.cpu()[rows], g["y_rows"], "cora: y rows")
    assert_close(grads["x"].cpu()[rows], g["g_x_rows"], "cora: g_x rows")
    for k in params:
        assert_close(grads[k], g["grad." + k], f"cora: grad {k}")
    assert y.double().sum().item() == pytest.approx(float(g["y_sum64"]), rel=1e-6, abs=1e-3)


BASE = dict(non_linearity="relu", normalization="none", learnable_scalar=True, skip_connection=True, normalize=True)


def _oracle_parity(ei, n, d, r, seed, tag, three_d=False, **over):
    kw = dict(BASE, **over)
    x, g_out, params = make_inputs(n, d, r, seed=seed)
    ours = run_ours(x, ei, params, g_out, three_d=three_d, **kw)
    mask = None
    if kw["non_linearity"] == "relu":
        mask = (impl.LAST_SAVED["zp"][:, :r] > 0).cpu()       # feed the oracle our ReLU decisions ...
    ref = run_oracle(x, ei, params, g_out, mask=mask, **kw)
    if mask is not None:                                      # ... after checking they only differ at ~0
        h1 = ref[2].last_preact
        diff = mask != (h1 > 0)
        assert diff.sum().item() <= 1e-5 * mask.numel() + 2
        if diff.any():
            assert h1[diff].abs().max().item() <= 1e-5 * h1.abs().max().item()
    compare(ours, ref, tag, g_out=g_out, batchnorm=kw["normalization"] == "batch_norm")
    return ours, ref


def test_config1_molecule_batch_gin_width():
    """configs[1] shapes: batch of 32 molecules (~25 nodes each), hidden 300, rank 16."""
    ei, _, n = molecule_batch(seed=0)
    _oracle_parity(ei, n, 300, 16, seed=1, tag="molecules")
    _oracle_parity(ei, n, 300, 16, seed=1, tag="molecules unnormalized", normalize=False)


@pytest.mark.parametrize("d", [32, 64])
@pytest.mark.parametrize("three_d", [False, True])
def test_config2_pubmed_shaped(d, three_d):
    """configs[2]: 19,717 nodes, 88,648 edges + pre-inserted self loops, NodeFormer passes [1,N,H]."""
    ei, n = make_graph("pubmed", seed=0)
    _oracle_parity(ei, n, d, 16, seed=2, tag=f"pubmed d={d}", three_d=three_d)


def test_config3_arxiv_shaped_full_size():
    """configs[3]: 169,343 nodes, 1,166,243 edges, hidden 256, rank 16 (full size; oracle ~ seconds)."""
    ei, n = make_graph("arxiv", seed=0)
    _oracle_parity(ei, n, 256, 16, seed=3, tag="arxiv")


def test_config3_arxiv_power_law_quarter_scale():
    ei, n = make_graph("arxiv", seed=1, power_law=True, scale=0.25)
    _oracle_parity(ei, n, 256, 16, seed=4, tag="arxiv power law")


def test_hub_rows_take_the_work_item_path():
    """A star (hub degree 6000 > kHubDeg = 512, cut into 512-neighbour work items) plus a random graph, directed
    extra edges into a second hub: exercises k_hub_partials and the ordered partial sums in all four hops."""
    n = 6001
    leaves = torch.arange(1, n, dtype=torch.int64)
    hub = torch.zeros(n - 1, dtype=torch.int64)
    rnd = symmetric_random_graph(n, 20000, seed=17)
    into7 = torch.randint(1, n, (1500,), generator=torch.Generator().manual_seed(3))       # directed: only in-edges
    ei = torch.cat([torch.stack([hub, leaves]), torch.stack([leaves, hub]), rnd,
                    torch.stack([into7, torch.full_like(into7, 7)])], dim=1)
    _oracle_parity(ei, n, 64, 16, seed=21, tag="hub star")
    _oracle_parity(ei, n, 64, 32, seed=22, tag="hub star r=32", normalize=False)


def test_config4_products_shaped_scaled():
    """configs[4] shape (hidden 256, rank 32) at 1/64 scale; the full size is covered by properties."""
    ei, n = make_graph("products", seed=0, scale=1 / 64)
    _oracle_parity(ei, n, 256, 32, seed=5, tag="products/64")


@pytest.mark.parametrize("over", [
    dict(non_linearity="silu"), dict(non_linearity="none"), dict(skip_connection=False),
    dict(learnable_scalar=False), dict(normalize=False), dict(normalization="layer_norm"),
    dict(normalization="batch_norm"), dict(normalize=False, non_linearity="silu", skip_connection=False),
])
def test_constructor_variants(over):
    n = 5000
    ei = symmetric_random_graph(n, 30000, seed=11)
    _oracle_parity(ei, n, 64, 16, seed=6, tag=str(over), **over)


@pytest.mark.parametrize("d,r", [(8, 8), (300, 8), (64, 32), (128, 64), (30, 5), (50, 20), (1028, 16)])
def test_shapes_including_padded_ones(d, r):
    n = 1500
    ei = symmetric_random_graph(n, 9000, seed=12)
    _oracle_parity(ei, n, d, r, seed=7, tag=f"d={d} r={r}")


def test_edge_cases_empty_edges_isolated_nodes_single_row():
    for n, ei in [(7, torch.zeros(2, 0, dtype=torch.int64)), (1, torch.zeros(2, 0, dtype=torch.int64)),
                  (65, torch.tensor([[0, 64], [64, 0]], dtype=torch.int64))]:
        for normalize in (True, False):
            _oracle_parity(ei, n, 16, 8, seed=8, tag=f"n={n} normalize={normalize}", normalize=normalize)


def test_non_contiguous_input_and_no_input_grad():
    n, d, r = 2000, 64, 16
    ei = symmetric_random_graph(n, 12000, seed=13)
    x, g_out, params = make_inputs(n, d, r, seed=9)
    wide = torch.randn(n, 2 * d)
    wide[:, :d] = x
    m = GConvAdapter(d, r, learnable_scalar=True)
    load_module_params(m, params)
    m = m.cuda()
    xs = wide.cuda()[:, :d]                       # row pitch 2d, inner stride 1
    assert not xs.is_contiguous()
    y1 = m(xs, ei.cuda())
    y2 = m(x.cuda(), ei.cuda())
    assert torch.equal(y1, y2)
    y1.backward(g_out.cuda())                     # x does not require grad: gX is skipped
    ref = run_oracle(x, ei, params, g_out, learnable_scalar=True)
    for k, p in m.named_parameters():
        assert_close(p.grad, ref[1][k], f"no-input-grad: {k}")


def test_bitwise_determinism_and_graph_cache_reuse():
    ei, n = make_graph("arxiv", seed=0, scale=0.5)
    x, g_out, params = make_inputs(n, 256, 16, seed=10)
    a = run_ours(x, ei, params, g_out, **BASE)
    b = run_ours(x, ei, params, g_out, **BASE)
    assert torch.equal(a[0], b[0])
    for k in a[1]:
        assert torch.equal(a[1][k], b[1][k]), k
    m = a[2]
    eic = ei.cuda()
    m(x.cuda(), eic); m(x.cuda(), eic)
    # run_ours built the structure once for its own device copy of edge_index; eic is a new tensor
    assert m.graph_cache.misses == 2 and m.graph_cache.hits == 1


@pytest.mark.parametrize("name,r", [("arxiv", 16), ("products", 32)])
def test_full_size_properties_linearity_and_adjoint(name, r):
    """Size-independent checks at BASELINE.json's full sizes (no oracle): with the identity
    non-linearity the adapter is affine in x, so (i) f(a x1 + b x2) - f(0) is linear and (ii) the
    backward is the exact adjoint: <g, J v> == <J^T g, v>."""
    ei, n = make_graph(name, seed=0)
    d = 256
    _, _, params = make_inputs(8, d, r, seed=11)
    m = GConvAdapter(d, r, non_linearity="none", learnable_scalar=True)
    load_module_params(m, params)
    m = m.cuda()
    eic = ei.cuda()
    gen = torch.Generator(device="cuda").manual_seed(0)
    x1 = torch.randn(n, d, device="cuda", generator=gen)
    x2 = torch.randn(n, d, device="cuda", generator=gen)
    g = torch.randn(n, d, device="cuda", generator=gen)
    with torch.no_grad():
        f0 = m(torch.zeros_like(x1), eic)
        f1, f2 = m(x1, eic) - f0, m(x2, eic) - f0
        f12 = m(0.5 * x1 - 2.0 * x2, eic) - f0
    assert_close(f12, 0.5 * f1 - 2.0 * f2, f"{name}: linearity", rtol=1e-4, atol_scale=1e-5)
    v = x1.clone().requires_grad_(True)
    y = m(v, eic)
    y.backward(g)
    lhs = (g.double() * f1.double()).sum().item()            # <g, J x1>
    rhs = (v.grad.double() * x1.double()).sum().item()       # <J^T g, x1>
    assert lhs == pytest.approx(rhs, rel=1e-5)
    del f0, f1, f2, f12, y
    torch.cuda.empty_cache()
