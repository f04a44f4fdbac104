"""Test double for the phase backend of gconv_adapter_b200.partition: the SAME phase formulas as
include/gca.h, written with dense torch ops on CPU over oracle/csr_ref.py structures.  It exists so the
multi-rank orchestration (row blocks, gather buffers, collectives) can run under gloo on a CPU box, and
so the r-wide reordering itself is checked against the PyG-order oracle without any CUDA code.  Test
infrastructure only - the product never imports it."""
import numpy as np
import torch

from oracle import csr_ref


class _Graph:
    def __init__(self, ei, n, normalize, lo, hi):
        b = csr_ref.build(ei.cpu().numpy(), n, normalize, lo, hi)
        self.lo, self.hi = lo, hi
        self.dis = torch.from_numpy(b["dis"][lo:hi].copy())
        self.rows = torch.from_numpy(np.repeat(np.arange(hi - lo), np.diff(b["rowptr"]))).long()
        self.cols = torch.from_numpy(b["colidx"].astype(np.int64))
        self.rows_t = torch.from_numpy(np.repeat(np.arange(hi - lo), np.diff(b["rowptr_t"]))).long()
        self.cols_t = torch.from_numpy(b["colidx_t"].astype(np.int64))

    def agg(self, full, transpose=False):
        rows, cols = (self.rows_t, self.cols_t) if transpose else (self.rows, self.cols)
        out = torch.zeros(self.hi - self.lo, full.shape[1], dtype=full.dtype)
        return out.index_add_(0, rows, full[cols])


def _act(h, act):
    return torch.relu(h) if act == 1 else (torch.nn.functional.silu(h) if act == 2 else h)


class CpuPhases:
    def build_graph(self, edge_index, num_nodes, normalize, lo, hi):
        return _Graph(edge_index, num_nodes, normalize, lo, hi)

    def fwd_project(self, g, x, wd, out_local, push=None):
        out_local.copy_(g.dis[:, None] * (x @ wd.t()))

    def fwd_hop1(self, g, p_full, bd, act, z_local, h1_local, hub=None, push=None):
        h = g.dis[:, None] * g.agg(p_full) + bd
        if h1_local is not None:
            h1_local[:h.shape[0]].copy_(h)
        z_local.copy_(g.dis[:, None] * _act(h, act))

    def fwd_hop2_up(self, g, z_full, x, wu, bu, scalar, skip, h2_local, y, hub=None):
        h2 = g.dis[:, None] * g.agg(z_full)
        h2_local[:h2.shape[0]].copy_(h2)
        s = scalar if scalar is not None else torch.ones(1)
        y.copy_(s * (h2 @ wu.t() + bu + (x if skip else 0)))

    def bwd_scratch(self, d, r, device):
        return {}

    def bwd_up(self, g, gy, h2_local, wu, scalar, gh2_local, scratch, push=None):
        s = scalar if scalar is not None else torch.ones(1)
        n = gy.shape[0]
        gh2_local.copy_(g.dis[:, None] * s * (gy @ wu))
        scratch["gu"] = gy.t() @ h2_local[:n]          # [d, r], unscaled
        scratch["col"] = gy.sum(0)

    def bwd_up_project(self, g, gy, wu, scalar, gh2_local, scratch):
        s = scalar if scalar is not None else torch.ones(1)
        gh2_local.copy_(g.dis[:, None] * s * (gy @ wu))

    def bwd_up_wgrad(self, g, gy, h2_local, scratch, d, r):
        scratch["gu"] = gy.t() @ h2_local[:gy.shape[0]]
        scratch["col"] = gy.sum(0)

    def bwd_hop2(self, g, gh2_full, z_local, h1_local, act, gh1_local, scratch, hub=None, push=None):
        gz = g.dis[:, None] * g.agg(gh2_full, transpose=True)
        n = gz.shape[0]
        if act == 1:
            gz = gz * (z_local > 0)
        elif act == 2:
            h = h1_local[:n]
            sg = torch.sigmoid(h)
            gz = gz * sg * (1 + h * (1 - sg))
        scratch["bd"] = gz.sum(0)
        gh1_local.copy_(g.dis[:, None] * gz)

    def bwd_hop1_down(self, g, gh1_full, x, gy, wd, scalar, skip, gp_local, gx, scratch, hub=None):
        s = scalar if scalar is not None else torch.ones(1)
        gp = g.dis[:, None] * g.agg(gh1_full, transpose=True)
        if gx is not None:
            gx.copy_(gp @ wd + (s * gy if skip else 0))
        scratch["gd"] = gp.t() @ x
        scratch["dot"] = (gy * x).sum() if (skip and scalar is not None) else torch.zeros(())

    def bwd_finalize(self, scratch, wu, bu, scalar, skip, g_wd, g_bd, g_wu, g_bu, g_s):
        s = scalar if scalar is not None else torch.ones(1)
        g_wd.copy_(scratch["gd"])
        g_bd.copy_(scratch["bd"])
        g_wu.copy_(s * scratch["gu"])
        g_bu.copy_(s * scratch["col"])
        if g_s is not None:
            g_s.copy_((scratch["dot"] + (scratch["gu"] * wu).sum() + (scratch["col"] * bu).sum()).reshape(1))
