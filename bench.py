#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of the PyG path

Workload at N=1 (BASELINE.json configs[2], the configuration the metric is quoted on): the SGCN
network of util/networks.py with the GCNConv branch (13 conv blocks, widths
4-16-32-64-128-256-256-512-256-256-128-64-32-16 + Linear(16,3)), fp32, forward + the reference's
two step losses + backward + Adam step, on a synthetic geodesic icosphere of frequency 316
(998 562 vertices, 5 991 360 directed edges).  One "step" = one such train step on one mask
(sgcn.py:129-146).  metric = directed edges * conv layers / second.

N > 1: one process per GPU (torchrun), each rank trains its own independent mesh of the same size
(config 5 style: self-prior = one model per mesh, no data-path collective) -> weak scaling;
timing is barrier + max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "sgcn_gcnconv_train_edge_layers_per_s"
UNIT = "edges/s"
N_LAYERS = 13
K1 = 4.0   # sgcn.py: loss = loss_p + k1 * loss_n, default k1 = 4.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--freq", type=int, default=0, help="icosphere frequency n (V = 10 n^2 + 2); 0 = 316 (1M vertices) for replicas, "
                                                        "632 (4M vertices, fits one GPU) for partition")
    ap.add_argument("--conv", default="gcnconv", choices=["gcnconv", "chebconv"])
    ap.add_argument("--cpu-freq", type=int, default=0, help="icosphere frequency of the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-kernel CUDA-event roofline pass")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="replay the whole train step as one CUDA graph (semigcn_b200/graphed.py): the launch-bound small-mesh regime")
    ap.add_argument("--order", default="given", choices=["given", "morton"],
                    help="partition mode: vertex numbering the contiguous cut is taken on (morton: Z-order renumbering, balanced halos; "
                         "CPU-tested, not yet timed on GPUs)")
    ap.add_argument("--mode", default="replicas", choices=["replicas", "partition"],
                    help="replicas: one independent mesh per GPU (default, configs[2]/[4]); partition: ONE mesh vertex-partitioned over the "
                         "GPUs with per-propagation halo exchange over NCCL (configs[3], strong scaling)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json; sustained bf16 figure: kernels timed inside a long step)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of ``kernel`` (family_shape name) from the committed
    ``ncu --set full`` capture summarised in profiles/ncu_traffic.json (written by tools/ncu_traffic.py); None if that
    kernel shape has not been captured."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p)).get(kernel)
        return (float(d["dram_bytes"]), d.get("source")) if d else (None, None)
    except (OSError, ValueError, KeyError):
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def algorithmic_bytes_per_step(n: int, nnz: int, widths):
    """per GCNConv layer fwd = B_gemm_io + B_spmm(width the SpMM runs at), bwd = SpMM + dX + dW (+db)."""
    total = 0.0
    for i in range(len(widths) - 2):
        cin, cout = widths[i], widths[i + 1]
        cs = min(cin, cout)
        b_spmm = 4.0 * (2 * n * cs + (nnz + n) + (n + 1) + n)
        b_gemm = 4.0 * (n * (cin + cout) + cin * cout)
        total += (b_gemm + b_spmm) + (b_spmm + 4.0 * n * (cin + 2 * cout) + 4.0 * n * cin + 4.0 * n * cout)
    return total


def gcnconv_layer_bench(edge_index, n: int, nnz: int, dev, pk, cin: int = 256, cout: int = 256, reps: int = 10):
    """BASELINE.json's first metric, "GCNConv fwd+bwd edges/s & HBM GB/s fraction", on ONE drop-in layer (no BatchNorm
    around it): forward + backward (dX, dW, db) of ``GCNConv(cin, cout)`` on the bench mesh, CUDA events, inputs >> L2.
    Algorithmic bytes per SURVEY.md §8(d):  fwd = B_gemm_io + B_spmm,  bwd = B_spmm + N (Cin + 2 Cout) 4 + N Cin 4 + N Cout 4."""
    from semigcn_b200.nn import GCNConv
    torch.manual_seed(314)
    conv = GCNConv(cin, cout).to(dev)
    x = torch.randn(n, cin, device=dev, requires_grad=True)
    g = torch.randn(n, cout, device=dev)

    def step():
        conv.zero_grad(set_to_none=True)
        x.grad = None
        conv(x, edge_index).backward(g)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    cs = min(cin, cout)
    b_spmm = 4.0 * (2 * n * cs + (nnz + n) + (n + 1) + n)
    b_fwd = 4.0 * (n * (cin + cout) + cin * cout) + b_spmm
    b_bwd = b_spmm + 4.0 * n * (cin + 2 * cout) + 4.0 * n * cin + 4.0 * n * cout
    gbs = (b_fwd + b_bwd) / (ms / 1e3) / 1e9
    return {"layer": f"GCNConv({cin}, {cout}) forward + backward (dX, dW, db), {n} vertices, {nnz} directed edges",
            "ms": ms, "edges_per_s": nnz / (ms / 1e3), "algorithmic_GB": (b_fwd + b_bwd) / 1e9, "GBps": gbs,
            "hbm_frac": gbs / pk["hbm_gbs"], "reps": reps}


def make_problem(freq: int, device, seed: int = 314):
    from semigcn_b200 import meshgen
    mesh = meshgen.icosphere(freq, device=device, dtype=torch.float64)
    g = torch.Generator(device="cpu").manual_seed(seed)
    nv = mesh.num_vertices
    bump = torch.randn(nv, 1, generator=g, dtype=torch.float64).to(device)
    ini = mesh.vs * (1.0 + 0.02 * bump)
    smo = meshgen.uniform_laplacian_smooth(ini, mesh.edge_index, 30)
    v_mask = torch.rand(nv, generator=g).to(device) > 0.03
    f_mask = v_mask[mesh.faces].all(dim=1)
    dms = (torch.rand(nv, 8, generator=g) > 0.1).float().to(device)
    fn = meshgen.face_normals(ini, mesh.faces)
    return dict(mesh=mesh, ini=ini, z1=(ini - smo).float(), x_pos=smo.float(), v_mask=v_mask, f_mask=f_mask, dms=dms, fn=fn)


def step_losses(out, prob):
    """The reference's step losses (sgcn.py:130-137), fp64 targets as in sgcn.py:127."""
    from semigcn_b200 import losses
    return losses.sgcn_step_loss(out, prob["mesh"].faces, prob["ini"], prob["fn"], prob["v_mask"], prob["f_mask"], K1)


# ------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    from semigcn_b200 import _lib, ops, profile
    from semigcn_b200.data import Data
    from semigcn_b200.networks import SingleScaleGCN, SGCN_WIDTHS
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    _lib.load()
    prob = make_problem(args.freq, dev, seed=314 + rank)
    mesh = prob["mesh"]
    n, nnz = mesh.num_vertices, mesh.nnz
    torch.manual_seed(314)
    net = SingleScaleGCN(dev, conv=args.conv).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=0.01, capturable=args.cuda_graph)
    data = Data(z1=prob["z1"], x_pos=prob["x_pos"], edge_index=mesh.edge_index)
    graphed = None
    if args.cuda_graph:
        from semigcn_b200.graphed import GraphedTrainStep
        graphed = GraphedTrainStep(net, lambda out: step_losses(out, prob), opt, prob["z1"], prob["x_pos"], mesh.edge_index,
                                   prob["dms"][:, 0:1].contiguous())

    def one_step(i, data_in, dm):
        if graphed is not None:      # inputs (device or pinned host) are copied into the graph's static buffers, then one replay
            return graphed(dm, None if data_in is data else data_in.z1, None if data_in is data else data_in.x_pos)
        opt.zero_grad(set_to_none=True)
        out = net(data_in, dm)
        loss = step_losses(out, prob)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ("value")
    for i in range(args.warmup):
        one_step(i, data, prob["dms"][:, i % 8:i % 8 + 1])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        one_step(i, data, prob["dms"][:, i % 8:i % 8 + 1])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.LAUNCHES - launches0
    clocks = sampler.stop()

    # ---------------- per-kernel-family roofline pass (CUDA events on the launching stream)
    fam = {}
    if not args.no_profile and rank == 0 and graphed is None:
        with profile.KernelProfile() as kp:
            for i in range(min(args.steps, 5)):
                one_step(i, data, prob["dms"][:, i % 8:i % 8 + 1])
            fam = kp.summary()
            prof_steps = min(args.steps, 5)

    # ---------------- end-to-end timing through the public API with HOST buffers ("e2e")
    h = {k: prob[k].cpu().pin_memory() for k in ("z1", "x_pos")}
    h["edge_index"] = mesh.edge_index.cpu().pin_memory()
    h["dms"] = [prob["dms"][:, j:j + 1].contiguous().cpu().pin_memory() for j in range(8)]
    d_z1, d_xp, d_ei = torch.empty_like(prob["z1"]), torch.empty_like(prob["x_pos"]), torch.empty_like(mesh.edge_index)
    d_dm = torch.empty((n, 1), dtype=torch.float32, device=dev)
    loss_host = torch.empty((), dtype=torch.float64).pin_memory()

    def e2e_step(i):
        if graphed is not None:      # graph mode: the mesh (edge_index) is fixed at capture; z1 / x_pos / mask come from pinned host memory
            loss = one_step(i, Data(z1=h["z1"], x_pos=h["x_pos"], edge_index=None), h["dms"][i % 8])
            loss_host.copy_(loss.detach(), non_blocking=True)
            return
        # util/networks.py:65 uploads z1 / x_pos / edge_index on every forward, :77 the mask
        d_z1.copy_(h["z1"], non_blocking=True)
        d_xp.copy_(h["x_pos"], non_blocking=True)
        d_ei.copy_(h["edge_index"], non_blocking=True)
        d_dm.copy_(h["dms"][i % 8], non_blocking=True)
        loss = one_step(i, Data(z1=d_z1, x_pos=d_xp, edge_index=d_ei), d_dm)
        loss_host.copy_(loss.detach(), non_blocking=True)

    for i in range(max(1, args.warmup)):
        e2e_step(i)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        e2e_step(i)
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    h2d = sum(t.numel() * t.element_size() for t in ((h["z1"], h["x_pos"], h["dms"][0]) if graphed is not None else
                                                       (h["z1"], h["x_pos"], h["edge_index"], h["dms"][0])))

    # ---------------- max over ranks
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return None
    units = float(nnz) * N_LAYERS * args.steps * world
    line = {
        "metric": METRIC, "value": units / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"SGCN ({args.conv}, 13 blocks, widths {SGCN_WIDTHS}) fwd + step losses + bwd + Adam on icosphere n={args.freq} "
                               f"({n} vertices, {nnz} directed edges) per GPU; BASELINE.json configs[2]",
                   "vertices": n, "directed_edges": nnz, "conv_layers": N_LAYERS, "parallelism": f"replicas x{world} (independent meshes, no collective)",
                   "l2_policy": "inputs larger than L2 (activations of one step ~16 GB >> 126 MB); no explicit flush",
                   "e2e_inputs": "z1, x_pos, edge_index, mask uploaded from pinned host memory every step (as util/networks.py:65,77); "
                                 "edge_index re-upload forces a CSR rebuild every step"},
        "train_steps_per_s": args.steps * world / (ms / 1e3),
        "e2e": {"value": units / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if graphed is not None:
        line["config"]["cuda_graph"] = "whole train step replayed as one CUDA graph (semigcn_b200/graphed.py); gpu_launches counted at capture"
        line["config"]["e2e_inputs"] = "z1, x_pos, mask uploaded from pinned host memory every step into the graph's static buffers; edge_index fixed at capture"
        line["gpu_launches"] = graphed.launches_per_replay * args.steps
    pk = peaks()
    if fam:
        # families (gemm, spmm, ...) for the step breakdown; the roofline block describes the DOMINANT KERNEL = the one
        # (family, shape) with the most time per step: its launches are identical, so "per launch" means one thing
        groups = {}
        for name, a in fam.items():
            key = name.split("_c")[0].split("_n")[0]
            gsum = groups.setdefault(key, dict(ms=0.0, bytes=0.0, flops=0.0, launches=0))
            for k2 in ("ms", "bytes", "flops", "launches"):
                gsum[k2] += a[k2]
        tot_ms = sum(v["ms"] for v in groups.values())
        dom = max(fam, key=lambda k: fam[k]["ms"])
        d = fam[dom]
        gbs = d["bytes"] / (d["ms"] / 1e3) / 1e9
        tfs = d["flops"] / (d["ms"] / 1e3) / 1e12
        t_hbm, t_tc = d["bytes"] / (pk["hbm_gbs"] * 1e9), d["flops"] / (pk["bf16_tflops"] * 1e12)
        if t_hbm >= t_tc:
            roof = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"]}
        else:
            roof = {"bound": "tensor", "achieved": tfs, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": tfs / pk["bf16_tflops"]}
        traffic, traffic_src = ncu_traffic(dom)
        roof.update({"traffic": traffic, "kernel": dom, "launches_per_step": d["launches"] / prof_steps,
                     "avg_launch_ms": d["ms"] / d["launches"], "algorithmic_bytes_per_launch": d["bytes"] / d["launches"],
                     "share_of_step_kernel_time": d["ms"] / tot_ms,
                     "peak_source": pk["source"], "traffic_source": traffic_src})
        line["roofline"] = roof
        line["kernel_families"] = {k: {"ms_per_step": v["ms"] / prof_steps, "GBps": v["bytes"] / (v["ms"] / 1e3) / 1e9,
                                       "TFLOPs": v["flops"] / (v["ms"] / 1e3) / 1e12, "launches_per_step": v["launches"] / prof_steps}
                                   for k, v in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])}
        line["kernel_shapes_top"] = {k: {"ms_per_step": v["ms"] / prof_steps, "GBps": v["bytes"] / (v["ms"] / 1e3) / 1e9,
                                         "launches_per_step": v["launches"] / prof_steps}
                                     for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])[:8]}
        alg = algorithmic_bytes_per_step(n, nnz, SGCN_WIDTHS)
        line["step_algorithmic_GB"] = alg / 1e9
        line["step_hbm_frac"] = alg / (ms / args.steps / 1e3) / 1e9 / pk["hbm_gbs"]
    if not args.no_profile and graphed is None:
        try:            # an extra, self-contained measurement: it must never cost the bench line
            torch.cuda.empty_cache()
            line["gcnconv_layer"] = gcnconv_layer_bench(mesh.edge_index, n, nnz, dev, pk)
        except Exception as exc:   # noqa: BLE001
            line["gcnconv_layer"] = {"error": repr(exc)[:300]}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(args, steps=2, warmup=1, freq=args.cpu_freq or 100)
    return line


# ------------------------------------------------------------------------------------------
def cpu_reference(args, steps: int, warmup: int, freq: int):
    """The oracle (restatement of the PyG 2.2.0 CPU path -- PyG itself is not installable here)
    timed on the host cores on a bounded sample of the same workload."""
    from oracle import pyg_ref as O
    from semigcn_b200 import meshgen
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    prob = make_problem(freq, "cpu")
    mesh = prob["mesh"]
    torch.manual_seed(314)
    net = O.SingleScaleGCN(args.conv)
    opt = torch.optim.Adam(net.parameters(), lr=0.01)

    def step(i):
        opt.zero_grad(set_to_none=True)
        out = net(prob["z1"], prob["x_pos"], mesh.edge_index, prob["dms"][:, i % 8:i % 8 + 1])
        loss = O.mask_pos_rec_loss(out, prob["ini"], prob["v_mask"]) + \
            K1 * O.mask_norm_rec_loss(O.compute_fn(out, mesh.faces), prob["fn"], prob["f_mask"])
        loss.backward()
        opt.step()

    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    dt = time.perf_counter() - t0
    return {"value": mesh.nnz * N_LAYERS * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"same network/losses/optimizer on icosphere n={freq} ({mesh.num_vertices} vertices, {mesh.nnz} directed edges), "
                      f"{steps} steps after {warmup} warm-up, torch {torch.__version__} CPU, {cores} threads; "
                      "restatement of the PyG CPU path (oracle/pyg_ref.py), not PyG itself",
            "ms_per_step": dt / steps * 1e3}


def run_reference(args, rank, world):
    if rank != 0:
        return None
    from semigcn_b200.networks import SGCN_WIDTHS
    freq = args.cpu_freq or (100 if args.steps + args.warmup <= 16 else 50)
    cb = cpu_reference(args, steps=args.steps, warmup=args.warmup, freq=freq)
    return {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"SGCN ({args.conv}, 13 blocks, widths {SGCN_WIDTHS}) fwd + step losses + bwd + Adam; CPU sample: {cb['sample']}"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_partition(args, rank, world, local_rank):
    """BASELINE.json configs[3]: ONE mesh, vertices partitioned over the ranks (contiguous ranges of the generator's
    grid-like numbering), one halo exchange per propagation (all-to-all over NCCL), SyncBN partial all-gather,
    weight-gradient all-reduce.  Total work is fixed as N grows -> strong scaling.  The step is forward + the masked
    position loss (the face-normal loss needs faces across cuts; not partitioned yet) + backward + gradient sum + Adam."""
    from semigcn_b200 import _lib, partition
    from semigcn_b200._lib import MODE_CHEB, MODE_GCN
    from semigcn_b200.data import Data
    from semigcn_b200.dist import TorchComm, dist_mask_pos_rec_loss, register_partition, sync_gradients
    from semigcn_b200.networks import SingleScaleGCN, SGCN_WIDTHS
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    _lib.load()
    comm = TorchComm()
    prob = make_problem(args.freq, dev, seed=314)            # the same mesh on every rank
    mesh = prob["mesh"]
    n, nnz = mesh.num_vertices, mesh.nnz
    ei_global = mesh.edge_index
    if args.order == "morton":
        perm = partition.morton_order(mesh.vs)
        ei_global, z1_, xp_, ini_, vm_, dms_ = partition.renumber(perm, mesh.edge_index, prob["z1"], prob["x_pos"], prob["ini"],
                                                                  prob["v_mask"], prob["dms"])
        prob.update(z1=z1_, x_pos=xp_, ini=ini_, v_mask=vm_, dms=dms_)
        ei_global = ei_global.contiguous()
    plan = partition.build_plan(ei_global, n, rank, world)
    lo, hi = plan.lo, plan.hi
    own = {k: prob[k][lo:hi].contiguous() for k in ("z1", "x_pos", "ini", "v_mask", "dms")}
    halo_rows = plan.n_ghost
    del prob, mesh, ei_global
    torch.cuda.empty_cache()
    ei_local = register_partition(plan, comm, modes=(MODE_GCN if args.conv == "gcnconv" else MODE_CHEB,))
    torch.manual_seed(314)
    net = SingleScaleGCN(dev, conv=args.conv).to(dev)
    net.comm = comm
    opt = torch.optim.Adam(net.parameters(), lr=0.01)
    data = Data(z1=own["z1"], x_pos=own["x_pos"], edge_index=ei_local)

    def one_step(i):
        opt.zero_grad(set_to_none=True)
        out = net(data, own["dms"][:, i % 8:i % 8 + 1])
        loss = dist_mask_pos_rec_loss(out, own["ini"], own["v_mask"], comm)
        loss.backward()
        sync_gradients(net, comm)
        opt.step()
        return loss

    def barrier():
        torch.distributed.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        one_step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = one_step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.LAUNCHES - launches0
    clocks = sampler.stop()
    loss_host = torch.empty((), dtype=torch.float64).pin_memory()
    # e2e: this rank's slice of z1 / x_pos / mask uploaded from pinned host memory every step, loss read back
    h = {k: own[k].cpu().pin_memory() for k in ("z1", "x_pos")}
    h_dm = [own["dms"][:, j:j + 1].contiguous().cpu().pin_memory() for j in range(8)]
    d_dm = torch.empty((hi - lo, 1), dtype=torch.float32, device=dev)

    def e2e_step(i):
        own["z1"].copy_(h["z1"], non_blocking=True)
        own["x_pos"].copy_(h["x_pos"], non_blocking=True)
        d_dm.copy_(h_dm[i % 8], non_blocking=True)
        opt.zero_grad(set_to_none=True)
        out = net(data, d_dm)
        loss = dist_mask_pos_rec_loss(out, own["ini"], own["v_mask"], comm)
        loss.backward()
        sync_gradients(net, comm)
        opt.step()
        loss_host.copy_(loss.detach(), non_blocking=True)

    e2e_step(0)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        e2e_step(i)
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    h2d = sum(t.numel() * t.element_size() for t in (h["z1"], h["x_pos"], h_dm[0]))
    t = torch.tensor([ms, ms_e2e, float(halo_rows), float(h2d)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
    ms, ms_e2e = float(tmax[0]), float(tmax[1])
    if rank != 0:
        return None
    units = float(nnz) * N_LAYERS * args.steps
    return {
        "metric": METRIC, "value": units / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"SGCN ({args.conv}, 13 blocks, widths {SGCN_WIDTHS}) fwd + masked position loss + bwd + gradient all-reduce + Adam "
                               f"on ONE icosphere n={args.freq} ({n} vertices, {nnz} directed edges) vertex-partitioned over {world} GPU(s); "
                               "BASELINE.json configs[3] at the largest size that also fits one GPU",
                   "vertices": n, "directed_edges": nnz, "conv_layers": N_LAYERS,
                   "parallelism": f"vertex partition x{world}: halo all-to-all per propagation (NCCL), SyncBN partial all-gather, grad all-reduce",
                   "halo_rows_total": int(t[2]), "halo_rows_max_per_rank": int(tmax[2]), "vertex_order": args.order,
                   "l2_policy": "inputs larger than L2; no explicit flush"},
        "train_steps_per_s": args.steps / (ms / 1e3),
        "e2e": {"value": units / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(t[3]), "d2h_bytes_per_step": 8 * world,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks, "final_loss": float(loss.detach()),
    }


def main():
    args = parse()
    if args.freq == 0:
        args.freq = 632 if args.mode == "partition" else 316
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        line = run_reference(args, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return
    from semigcn_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        # the built .so normally travels with the tree; a bare checkout builds it (nvcc, ~1 min; local rank 0 only)
        if local_rank == 0:
            _lib.build_library(verbose=False)
        else:
            for _ in range(600):
                if os.path.exists(_lib.LIB_PATH):
                    break
                time.sleep(0.5)
            time.sleep(2.0)
    dist_on = world > 1 or args.mode == "partition"
    if dist_on:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        os.environ.setdefault("RANK", "0")
        os.environ.setdefault("WORLD_SIZE", "1")
        torch.distributed.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    line = (run_partition if args.mode == "partition" else run_ours)(args, rank, world, local_rank)
    if line is not None:
        print(json.dumps(line), flush=True)
    if dist_on:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
