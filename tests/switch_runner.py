"""Child process of tests/test_gpu_switches.py: the run-time switches of libgca (GCA_*) are latched in statics at the
first use, so every switch setting gets its own interpreter.  Runs the oracle parity of the adapter forward + backward
for the named case and prints ONE JSON line: the kernel variant libgca picked for every phase (gca_profile_report) -
so the parent can assert that the switch really selected the code path it names - and "ok"."""
import ctypes
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import torch  # noqa: E402

from gconv_adapter_b200 import _cabi  # noqa: E402
from gconv_adapter_b200.graphs.synthetic import make_graph  # noqa: E402
from test_gpu_adapter import _oracle_parity  # noqa: E402

CASES = {
    # name: (workload, scale, d, r, overrides)
    "arxiv_relu": ("arxiv", 0.25, 256, 16, {}),
    "arxiv_silu": ("arxiv", 0.25, 256, 16, {"non_linearity": "silu"}),
    "products64_relu": ("products", 1 / 64, 256, 32, {}),
    "products64_silu_noskip": ("products", 1 / 64, 256, 32, {"non_linearity": "silu", "skip_connection": False}),
    # small graphs: the fused one-kernel-per-direction path (gca_small.cu) unless GCA_DISABLE_SMALL=1
    "cora_relu": ("cora", 1.0, 64, 8, {}),
    "cora_silu_noskip": ("cora", 1.5, 64, 16, {"non_linearity": "silu", "skip_connection": False}),
    "molecules_relu": ("molecules", 1.0, 300, 16, {}),
    "small_r32_noscalar": ("cora", 1.0, 128, 32, {"learnable_scalar": False}),
    "small_r64_unnormalized": ("cora", 0.5, 64, 64, {"normalize": False}),
    # hub-heavy degree distribution: rows far longer than 512 neighbours, swept by whole warps in the fused kernels
    "small_powerlaw": ("cora", 1.0, 64, 16, {}, True),
}


def main():
    case = sys.argv[1]
    name, scale, d, r, over = CASES[case][:5]
    power_law = len(CASES[case]) > 5 and CASES[case][5]
    lib = _cabi.load()
    ei, n = make_graph(name, seed=0, scale=scale, power_law=power_law)
    lib.gca_profile_enable(1)
    _oracle_parity(ei, n, d, r, seed=31, tag=case, **over)      # raises on any parity failure
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.gca_profile_report(buf, len(buf))
    lib.gca_profile_enable(0)
    prof = json.loads(buf.value.decode())
    print(json.dumps({"ok": True, "case": case, "n": n, "max_in_degree": int(torch.bincount(ei[1], minlength=n).max()), "variants": {k: v["variant"] for k, v in prof.items()}}))


if __name__ == "__main__":
    main()
