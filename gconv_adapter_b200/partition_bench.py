"""Multi-GPU leg of bench.py: the workload graph row-partitioned over the ranks of one node (torchrun: one process per
GPU), fwd + bwd per step, device-timed, max over ranks.

* exchange: NVLink peer memory by default (gconv_adapter_b200/partition.py: producers push rows into the peers' gathered
  buffers, one-warp barrier kernels, rank-ordered peer all-reduce) - no NCCL call inside a step, so the step is captured
  in a CUDA graph and replayed (GCA_BENCH_EAGER=1: eager launches; GCA_PARTITION_COMM=nccl: torch.distributed collectives);
* self-checking: every rank draws its rows of X / gY from one seed, rank 0 runs the 1-GPU module on the FULL inputs once
  before the timed region, and the line carries ``parity_max_err_over_max_ref`` for Y, gX and every parameter gradient;
* the default (arxiv-shaped) run also measures the products-shaped graph (BASELINE.json configs[4]) at the same GPU count
  and reports it as the ``products`` sub-record, so the 1/2/4/8 curve of both partitioned configs is driver-observed.
"""
from __future__ import annotations

import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist


def _measure(name, args, lib, dev, rank, world, steps, warmup, with_e2e):
    from .graphs.synthetic import SHAPES, input_rows, make_graph, make_inputs
    from .partition import PartitionedGConvAdapter, row_block
    from . import GConvAdapter, GraphCache

    ei, n = make_graph(name, seed=0, power_law=args.power_law)       # every rank builds the same graph
    shp = SHAPES[name]
    d, r = shp.hidden, shp.rank
    e = ei.size(1)
    lo, hi = row_block(n, world, rank)
    seed = 1234
    xd, gd = input_rows(lo, hi, d, seed, dev)                        # only this rank's rows live on the device
    xd.requires_grad_(True)
    _, _, params = make_inputs(8, d, r, seed=0)
    m = PartitionedGConvAdapter(d, r, learnable_scalar=True)
    sd = m.state_dict()
    with torch.no_grad():
        for k, v in params.items():
            sd[k].copy_(v)
    m = m.to(dev)
    eid = ei.to(dev)
    del ei
    holder = {}

    def step():
        xd.grad = None
        for p in m.parameters():
            p.grad = None
        y = m(xd, eid, n)
        y.backward(gd)
        holder["y"] = y.detach()     # (a live y would keep the autograd graph - and its AccumulateGrad nodes - across steps)

    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    dist.barrier()

    # ---- parity against the 1-GPU module on the full inputs (rank 0), outside the timed region ----
    s_rows = (n + world - 1) // world
    pad = lambda t: torch.cat([t, t.new_zeros(s_rows - t.shape[0], t.shape[1])]) if t.shape[0] < s_rows else t
    y_all = [torch.empty(s_rows, d, device=dev) for _ in range(world)] if rank == 0 else None
    gx_all = [torch.empty(s_rows, d, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(pad(holder["y"]), y_all, dst=0)
    dist.gather(pad(xd.grad), gx_all, dst=0)
    parity = None
    if rank == 0:
        xf, gf = input_rows(0, n, d, seed, dev)
        xf.requires_grad_(True)
        ref = GConvAdapter(d, r, learnable_scalar=True)
        sd = ref.state_dict()
        with torch.no_grad():
            for k, v in params.items():
                sd[k].copy_(v)
        ref = ref.to(dev)
        ref.graph_cache = GraphCache()
        yr = ref(xf, eid)
        yr.backward(gf)
        torch.cuda.synchronize()
        rel = lambda a, b: ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()
        own = lambda parts: torch.cat([t[:min(s_rows, n - k * s_rows)] for k, t in enumerate(parts) if n - k * s_rows > 0])
        parity = {"y": rel(own(y_all), yr.detach()), "g_x": rel(own(gx_all), xf.grad)}
        for (k, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
            parity["g_" + k] = rel(p.grad, q.grad)
        # the scalar's gradient is ONE heavily cancelling sum of N d products: its fp32 noise floor relative to the value
        # (the tolerance tests/test_gpu_adapter.py::compare uses), for reading g_scalar against
        parity["g_scalar_fp32_noise_floor"] = (5e-7 * float((gf.abs().double() * yr.detach().abs().double()).sum())
                                               / abs(float(ref.scalar)) / max(abs(float(ref.scalar.grad)), 1e-30))
        del xf, gf, yr, ref, y_all, gx_all
        torch.cuda.empty_cache()
    dist.barrier()

    l0 = lib.gca_launch_count()
    step()
    per_step_launches = lib.gca_launch_count() - l0
    torch.cuda.synchronize()

    # ---- CUDA graph of one step (kernels, pushes, barrier kernels, peer all-reduce): PartitionedGConvAdapter.capture_step ----
    peer = getattr(m._comm, "fused_push", False)
    replay = m.capture_step(step) if os.environ.get("GCA_BENCH_EAGER", "0") != "1" else None
    run = replay if replay is not None else step

    from bench import ClockSampler   # noqa: PLC0415 - bench.py is the entry script
    sampler = ClockSampler(dev.index) if rank == 0 else None
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        run()
    t1.record()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = None
    if sampler:
        # keep the GPU under the same load for ~1 s so nvidia-smi (100 ms period) sees the clocks of this kernel mix;
        # every rank runs the same number of extra steps (the barrier kernels need all of them)
        for _ in range(max(1, int(1000.0 / max(ms.item(), 0.05)) // 4)):
            run()
        torch.cuda.synchronize()
        clocks = sampler.stop()
    else:
        for _ in range(max(1, int(1000.0 / max(ms.item(), 0.05)) // 4)):
            run()
        torch.cuda.synchronize()
    launches = torch.tensor([per_step_launches * steps], device=dev)
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)

    # ---- per-kernel device times on rank 0 (eager, profiled) for the roofline of the dominant kernel ----
    prof = None
    lib.gca_profile_enable(1)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.gca_profile_report(buf, len(buf))
    lib.gca_profile_enable(0)
    if rank == 0:
        prof = {k: v["ms"] / max(v["launches"], 1) for k, v in json.loads(buf.value.decode()).items()}
    dist.barrier()

    # ---- e2e: host rows -> device, fwd + bwd, Y and the loss back to the host, every step ----
    e2e = None
    if with_e2e:
        x_host = xd.detach().cpu().pin_memory()
        y_host = torch.empty_like(x_host).pin_memory()
        x_dev = torch.empty_like(xd)

        def e2e_step():
            for p in m.parameters():
                p.grad = None
            x_dev.copy_(x_host, non_blocking=True)
            xin = x_dev.detach().requires_grad_(True)
            y = m(xin, eid, n)
            loss = (y * gd).sum()
            loss.backward()
            y_host.copy_(y.detach(), non_blocking=True)
            return loss.item()

        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        dist.barrier()
        k2 = max(5, steps // 3)
        t0.record()
        for _ in range(k2):
            e2e_step()
        t1.record()
        torch.cuda.synchronize()
        ms2 = torch.tensor([t0.elapsed_time(t1) / k2], device=dev)
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        e2e = {"ms_per_step": round(ms2.item(), 4), "steps": k2, "h2d_bytes_per_step": 4 * n * d,
               "d2h_bytes_per_step": 4 * n * d + 4 * world}
    rec = {"name": name, "n": n, "e": e, "d": d, "r": r, "ms_per_step": ms.item(), "launches": int(launches.item()),
           "graph": replay is not None, "comm": "peer" if peer else "collective", "parity": parity, "clocks": clocks, "prof": prof,
           "e2e": e2e, "rows_per_rank": hi - lo}
    m.close()
    del m, eid, xd, gd
    torch.cuda.empty_cache()
    return rec


def run_partitioned(args, metric: str, unit: str) -> None:
    from . import _cabi
    from bench import algorithmic_bytes, peaks   # noqa: PLC0415

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()
    name = args.workload or "arxiv"
    main = _measure(name, args, lib, dev, rank, world, args.steps, args.warmup, with_e2e=not args.no_e2e)
    sub = None
    if args.workload is None and os.environ.get("GCA_BENCH_NO_PRODUCTS", "0") != "1":
        sub = _measure("products", args, lib, dev, rank, world, max(5, args.steps // 3), 3, with_e2e=False)

    if rank == 0:
        peak, peak_src = peaks()

        def roof(rec):
            """Whole-job roofline of a partitioned step: A_min of the full graph over (ms x world GPUs x peak)."""
            e_prime = rec["e"] + rec["n"]           # (upper bound: pre-existing self loops are replaced, not added)
            per, a_min = algorithmic_bytes(rec["n"], e_prime, rec["d"], rec["r"])
            gbs = a_min / 1e6 / rec["ms_per_step"]
            dom, dom_ms = None, 0.0
            for k, v in (rec["prof"] or {}).items():
                if v > dom_ms and k in per:
                    dom, dom_ms = k, v
            out = {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                   "step_alg_bytes": a_min, "step_achieved_all_gpus": round(gbs, 1),
                   "step_frac_of_measured_peak_per_gpu": round(gbs / world / peak, 4)}
            if dom:
                # rank 0's share of the rows
                share = rec["rows_per_rank"] / rec["n"]
                e_share = int(e_prime * share)
                per_r, _ = algorithmic_bytes(rec["rows_per_rank"], e_share, rec["d"], rec["r"])
                out.update({"kernel": dom, "achieved": round(per_r[dom] / 1e6 / dom_ms, 1), "frac": round(per_r[dom] / 1e6 / dom_ms / peak, 4),
                            "ms_per_launch": round(dom_ms, 5), "alg_bytes_per_launch": per_r[dom], "traffic": None,
                            "note": "rank 0, its rows only; eager profiled launches"})
            return out

        ms_step = main["ms_per_step"]
        nr = 4 * main["n"] * main["r"]
        line = {
            "metric": metric, "value": main["e"] / (ms_step / 1e3), "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{main['name']}-shaped: N={main['n']} E={main['e']} hidden={main['d']} rank={main['r']}",
                       "adapter": "relu, skip, learnable scalar, normalize=True", "graph": "static, structure cached",
                       "parallelism": f"row partition over {world} GPUs; per step 4 r-wide exchanges of [N, r] fp32 ({nr / 1e6:.1f} MB "
                                      f"each, (G-1)/G of it crosses NVLink) + 1 all-reduce of {2 * main['d'] * main['r'] + main['d'] + main['r'] + 1} floats",
                       "exchange": "NVLink peer memory: producers push rows, barrier kernels, peer all-reduce (no NCCL in the step)"
                                   if main["comm"] == "peer" else "torch.distributed (NCCL) all-gather / all-reduce",
                       "launch": "CUDA graph replay of one step" if main["graph"] else "eager",
                       "l2": "per-rank X/gY/Y/gX are %d MB each" % (4 * main["rows_per_rank"] * main["d"] // 1_000_000)},
            "parity_max_err_over_max_ref": main["parity"],
            "e2e": ({"value": main["e"] / (main["e2e"]["ms_per_step"] / 1e3), "unit": unit, **main["e2e"]} if main["e2e"] else None),
            "gpu_launches": main["launches"], "clocks": main["clocks"], "roofline": roof(main),
            "phases_rank0_ms": {k: round(v, 5) for k, v in (main["prof"] or {}).items()},
            "cpu_baseline": None,
        }
        if sub is not None:
            line["products"] = {
                "workload": f"products-shaped: N={sub['n']} E={sub['e']} hidden={sub['d']} rank={sub['r']}",
                "ms_per_step": sub["ms_per_step"], "value": sub["e"] / (sub["ms_per_step"] / 1e3), "unit": unit, "n_gpus": world,
                "launch": "CUDA graph replay of one step" if sub["graph"] else "eager", "exchange": sub["comm"],
                "parity_max_err_over_max_ref": sub["parity"], "roofline": roof(sub),
                "phases_rank0_ms": {k: round(v, 5) for k, v in (sub["prof"] or {}).items()},
            }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
