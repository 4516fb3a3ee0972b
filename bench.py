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

N > 1: one process per GPU (torchrun).  The headline `value` is the replica workload -- each rank trains its own
independent mesh of the same size (self-prior = one model per mesh, sgcn.py:78-80; no data-path collective) -> weak
scaling, barrier + max over ranks.  The SAME line then carries the runs the north star asks for next to it:
  "partition": ONE mesh vertex-partitioned over the N GPUs (BASELINE.json configs[3]): per-propagation halo all-to-all
               over NCCL overlapped with the interior rows, SyncBN, gradient all-reduce; strong scaling against the
               single-GPU step of the same mesh measured in the same invocation, plus the 16 M-vertex mesh at N >= 4;
  "batch64":   64 independent 100 k-vertex meshes spread over the N GPUs (configs[4]), several per GPU on streams,
               each as a whole-step CUDA graph.
At N = 1 the line also carries configs[0] / configs[1]: the SGCN and MGCN steps on a 10 242-vertex mesh (CUDA graph).
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
    ap.add_argument("--order", default="morton", choices=["given", "morton"],
                    help="partition mode: vertex numbering the contiguous cut is taken on (morton: Z-order patches, balanced halos; "
                         "given: the generator's numbering).  Both are followed by the interior-first order inside every rank's range")
    ap.add_argument("--mesh-order", default="given", choices=["given", "morton"],
                    help="replicas mode: numbering of the synthetic mesh handed to the network (given: the generator's row-by-row numbering; "
                         "morton: Z-order patches) -- an experiment knob for the locality of the aggregation kernel")
    ap.add_argument("--no-overlap", action="store_true", help="partition mode: do not overlap the halo exchange with the interior rows")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra blocks (partition / batch64 / configs[0], [1]) of the default line")
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
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  nvidia-smi takes a few hundred
    milliseconds to come up (longer on an 8-GPU box), so it is started BEFORE the warm-up steps; ``mark_start()`` / ``stop()``
    bracket the timed region and only the samples inside it are used -- unless the region was shorter than the sampling period,
    in which case the samples taken under the identical warm-up load are reported and ``window`` says so."""

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.t_start = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_start(self):
        self.t_start = time.time()

    def stop(self):
        t_end = time.time()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.06)                      # let the sample that covers the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = self.t_start if self.t_start is not None else 0.0
        inside = [ln for (t, ln) in self.lines if t0 <= t <= t_end + 0.06]
        window = "timed region"
        if not inside:                        # region shorter than the sampling period: the warm-up ran the same load
            inside = [ln for (_, ln) in self.lines[1:]] or [ln for (_, ln) in self.lines]
            window = "warm-up + timed region (timed region shorter than the sampling period)"
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in inside:
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
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm), "window": window}


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


def make_problem(freq: int, device, seed: int = 314, order: str = "given"):
    from semigcn_b200 import meshgen
    mesh = meshgen.icosphere(freq, device=device, dtype=torch.float64)
    if order == "morton":
        from semigcn_b200 import partition
        perm = partition.morton_order(mesh.vs)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(perm.numel(), device=perm.device)
        mesh = meshgen.SynthMesh(vs=mesh.vs[perm].contiguous(), faces=inv[mesh.faces].contiguous(), edges=inv[mesh.edges].contiguous(),
                                 edge_index=inv[mesh.edge_index].contiguous())
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
    prob = make_problem(args.freq, dev, seed=314 + rank, order=args.mesh_order)
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
        with profile.region("adam"):         # no-op unless the roofline pass below is active
            opt.step()
        return loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ("value")
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        one_step(i, data, prob["dms"][:, i % 8:i % 8 + 1])
    barrier()
    sampler.mark_start()
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
    loss_host = torch.empty((), dtype=torch.float64).pin_memory()

    from semigcn_b200.data import HostInputPipeline
    pipe = None
    if graphed is None:
        # util/networks.py:65 uploads z1 / x_pos / edge_index on every forward, :77 the mask.  Every step's inputs cross PCIe
        # from pinned host memory inside the timed region; the upload of step i + 1 runs on a copy stream while step i computes
        # (semigcn_b200.data.HostInputPipeline, two device slots).
        pipe = HostInputPipeline(dev, z1=h["z1"], x_pos=h["x_pos"], edge_index=h["edge_index"], dm=h["dms"][0])

    def submit(i):
        pipe.submit(z1=h["z1"], x_pos=h["x_pos"], edge_index=h["edge_index"], dm=h["dms"][i % 8])

    def e2e_step(i, last=False):
        if graphed is not None:      # graph mode: the mesh (edge_index) is fixed at capture; z1 / x_pos / mask come from pinned host memory
            loss = one_step(i, Data(z1=h["z1"], x_pos=h["x_pos"], edge_index=None), h["dms"][i % 8])
            loss_host.copy_(loss.detach(), non_blocking=True)
            return
        if not last:
            submit(i + 1)
        d = pipe.get()
        loss = one_step(i, Data(z1=d["z1"], x_pos=d["x_pos"], edge_index=d["edge_index"]), d["dm"])
        pipe.release()
        loss_host.copy_(loss.detach(), non_blocking=True)

    n_e2e_warm = max(1, args.warmup)
    if pipe is not None:
        submit(0)
    for i in range(n_e2e_warm):
        e2e_step(i)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(n_e2e_warm, n_e2e_warm + args.steps):
        e2e_step(i, last=(i == n_e2e_warm + args.steps - 1))
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
                   "e2e_inputs": "z1, x_pos, edge_index, mask uploaded from pinned host memory every step (as util/networks.py:65,77) through the "
                                 "public double-buffered input pipeline (semigcn_b200.data.HostInputPipeline: the upload of step i+1 overlaps step i); "
                                 "the re-uploaded edge_index is resolved by the CSR cache by content (no rebuild)"},
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
        line["kernel_families_sum_ms"] = tot_ms / prof_steps      # vs ms_per_step: what the per-family records account for
        alg = algorithmic_bytes_per_step(n, nnz, SGCN_WIDTHS)
        line["step_algorithmic_GB"] = alg / 1e9
        line["step_hbm_frac"] = alg / (ms / args.steps / 1e3) / 1e9 / pk["hbm_gbs"]
    if not args.no_profile and graphed is None:
        try:            # an extra, self-contained measurement: it must never cost the bench line
            torch.cuda.empty_cache()
            line["gcnconv_layer"] = gcnconv_layer_bench(mesh.edge_index, n, nnz, dev, pk)
        except Exception as exc:   # noqa: BLE001
            line["gcnconv_layer"] = {"error": repr(exc)[:300]}
    if not args.no_cpu_baseline and world == 1:        # the contract: rank 0 at N = 1 only
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
    # a FIXED sample whatever --steps / --warmup are: the 100 002-vertex icosphere (about 2 s per step on 16 host cores; the
    # 998 562-vertex mesh of our arm would take ~19 s per step, minutes for the driver's 25 steps).  edges/s normalises the size.
    freq = args.cpu_freq or 100
    cb = cpu_reference(args, steps=args.steps, warmup=args.warmup, freq=freq)
    return {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"SGCN ({args.conv}, 13 blocks, widths {SGCN_WIDTHS}) fwd + step losses + bwd + Adam; CPU sample: {cb['sample']}"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def _time_steps(step_fn, steps: int, warmup: int, barrier):
    """W untimed + K timed calls of ``step_fn(i)``, CUDA events on the current stream, barrier + synchronize on both sides."""
    for i in range(warmup):
        step_fn(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for i in range(steps):
        out = step_fn(i)
    e1.record()
    barrier()
    return e0.elapsed_time(e1), out


def single_gpu_measure(args, dev, freq: int, steps: int, warmup: int):
    """The UNPARTITIONED train step (forward + both step losses + backward + Adam, exactly the headline workload) on one
    icosphere of frequency ``freq`` on ``dev``: the like-for-like single-GPU rate the partitioned runs are compared with."""
    from semigcn_b200 import ops
    from semigcn_b200.data import Data
    from semigcn_b200.networks import SingleScaleGCN
    prob = make_problem(freq, dev, seed=314)
    mesh = prob["mesh"]
    torch.manual_seed(314)
    net = SingleScaleGCN(dev, conv=args.conv).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=0.01)
    data = Data(z1=prob["z1"], x_pos=prob["x_pos"], edge_index=mesh.edge_index)

    def one_step(i):
        opt.zero_grad(set_to_none=True)
        loss = step_losses(net(data, prob["dms"][:, i % 8:i % 8 + 1]), prob)
        loss.backward()
        opt.step()
        return loss

    ms, loss = _time_steps(one_step, steps, warmup, torch.cuda.synchronize)
    res = {"vertices": mesh.num_vertices, "directed_edges": mesh.nnz, "ms_per_step": ms / steps,
           "edges_per_s": float(mesh.nnz) * N_LAYERS * steps / (ms / 1e3), "final_loss": float(loss.detach())}
    del net, opt, data, prob, mesh
    ops.clear_graph_cache()
    torch.cuda.empty_cache()
    return res


def partition_measure(args, rank, world, local_rank, freq: int, steps: int, warmup: int, e2e: bool = True):
    """BASELINE.json configs[3]: ONE mesh, vertices partitioned over the ranks -- Morton patches (contiguous ranges of a
    Z-order renumbering), interior vertices first inside every range.  Per propagation one halo exchange (pack kernel +
    all-to-all over NCCL) that runs WHILE the interior rows are aggregated; SyncBN all-gathers one merged moment row per
    rank per layer; weight gradients are all-reduced once per step.  The step is the single-GPU step: forward + masked
    position loss + face-normal loss across the cuts + backward + gradient sum + Adam.  Total work is fixed as N grows
    -> strong scaling.  Returns a dict on rank 0 (None elsewhere); device-timed, max over ranks."""
    from semigcn_b200 import _lib, ops, partition
    from semigcn_b200._lib import MODE_CHEB, MODE_GCN
    from semigcn_b200.data import Data
    from semigcn_b200.dist import (TorchComm, dist_mask_norm_rec_loss, dist_mask_pos_rec_loss, register_partition, sync_gradients)
    from semigcn_b200.networks import SingleScaleGCN
    dev = torch.device(f"cuda:{local_rank}")
    comm = TorchComm()
    prob = make_problem(freq, dev, seed=314)            # the same mesh on every rank
    mesh = prob["mesh"]
    n, nnz = mesh.num_vertices, mesh.nnz
    ei, faces = mesh.edge_index, mesh.faces
    vt = [prob[k] for k in ("z1", "x_pos", "ini", "v_mask", "dms")]
    ranges = partition.vertex_ranges(n, world)

    def renumbered(perm, ei, faces, vt):
        ei, *vt = partition.renumber(perm, ei, *vt)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(n, device=dev)
        return ei, inv[faces], vt

    if args.order == "morton":
        ei, faces, vt = renumbered(partition.morton_order(mesh.vs), ei, faces, vt)
    ei, faces, vt = renumbered(partition.interior_first_order(ei, n, ranges), ei, faces, vt)
    ei = ei.contiguous()
    z1_, xp_, ini_, vm_, dms_ = vt
    plan = partition.build_plan(ei, n, rank, world, ranges=ranges)
    lo, hi = plan.lo, plan.hi
    own = {"z1": z1_[lo:hi].contiguous(), "x_pos": xp_[lo:hi].contiguous(), "ini": ini_[lo:hi].contiguous(),
           "v_mask": vm_[lo:hi].contiguous(), "dms": dms_[lo:hi].contiguous()}
    fid, f_loc = partition.local_faces(plan, faces)
    fn_loc, fm_loc = prob["fn"][fid].contiguous(), prob["f_mask"][fid].contiguous()
    halo_rows = plan.n_ghost
    del prob, mesh, ei, faces, vt, z1_, xp_, ini_, vm_, dms_, fid
    torch.cuda.empty_cache()
    mode = MODE_GCN if args.conv == "gcnconv" else MODE_CHEB
    ei_local = register_partition(plan, comm, modes=(mode,), overlap=not args.no_overlap)
    g_part = ops.graph_for(ei_local, hi - lo, mode)
    torch.manual_seed(314)
    net = SingleScaleGCN(dev, conv=args.conv).to(dev)
    net.comm = comm
    opt = torch.optim.Adam(net.parameters(), lr=0.01)
    data = Data(z1=own["z1"], x_pos=own["x_pos"], edge_index=ei_local)

    def loss_of(out):
        lp = dist_mask_pos_rec_loss(out, own["ini"], own["v_mask"], comm)
        ln = dist_mask_norm_rec_loss(out, g_part.halo, f_loc, fn_loc, fm_loc, comm)
        return lp + K1 * ln

    def one_step(i, dm=None):
        opt.zero_grad(set_to_none=True)
        loss = loss_of(net(data, own["dms"][:, i % 8:i % 8 + 1] if dm is None else dm))
        loss.backward()
        sync_gradients(net, comm)
        opt.step()
        return loss

    def barrier():
        torch.distributed.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(warmup):
        one_step(i)
    barrier()
    sampler.mark_start()
    launches0, bytes0, calls0 = _lib.LAUNCHES, g_part.halo.bytes_total, g_part.halo.calls
    ms, loss = _time_steps(one_step, steps, 0, barrier)
    launches = _lib.LAUNCHES - launches0
    halo_bytes = (g_part.halo.bytes_total - bytes0) / steps
    halo_calls = (g_part.halo.calls - calls0) / steps
    clocks = sampler.stop()
    ms_e2e, h2d = ms, 0
    if e2e:
        # this rank's slice of z1 / x_pos / mask uploaded from pinned host memory every step, loss read back
        loss_host = torch.empty((), dtype=torch.float64).pin_memory()
        h = {k: own[k].cpu().pin_memory() for k in ("z1", "x_pos")}
        h_dm = [own["dms"][:, j:j + 1].contiguous().cpu().pin_memory() for j in range(8)]
        d_dm = torch.empty((hi - lo, 1), dtype=torch.float32, device=dev)

        def e2e_step(i):
            own["z1"].copy_(h["z1"], non_blocking=True)
            own["x_pos"].copy_(h["x_pos"], non_blocking=True)
            d_dm.copy_(h_dm[i % 8], non_blocking=True)
            loss_host.copy_(one_step(i, d_dm).detach(), non_blocking=True)

        ms_e2e, _ = _time_steps(e2e_step, steps, 1, barrier)
        h2d = sum(t.numel() * t.element_size() for t in (h["z1"], h["x_pos"], h_dm[0]))
    t = torch.tensor([ms, ms_e2e, float(halo_rows), float(h2d), float(halo_bytes), float(g_part.n_interior), float(hi - lo)],
                     dtype=torch.float64, device=dev)
    tmax = t.clone()
    torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
    final_loss = float(loss.detach())
    del net, opt, data, own, g_part, plan
    ops.clear_graph_cache()
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    ms, ms_e2e = float(tmax[0]), float(tmax[1])
    units = float(nnz) * N_LAYERS * steps
    return {
        "vertices": n, "directed_edges": nnz, "n_gpus": world, "ms_per_step": ms / steps, "edges_per_s": units / (ms / 1e3),
        "e2e_ms_per_step": ms_e2e / steps, "e2e_edges_per_s": units / (ms_e2e / 1e3), "h2d_bytes_per_step": int(t[3]),
        "halo_rows_total": int(t[2]), "halo_rows_max_per_rank": int(tmax[2]), "interior_rows_frac": float(t[5]) / float(t[6]),
        "halo_exchanges_per_step": halo_calls, "halo_bytes_per_step_per_rank_max": float(tmax[4]), "vertex_order": args.order + " + interior-first",
        "overlap": not args.no_overlap, "gpu_launches": launches, "clocks": clocks, "final_loss": final_loss,
        "collectives": "per propagation 1 all_to_all_single (halo rows x C floats, overlapped with the interior rows); per BatchNorm layer 1 "
                       "all_gather_into_tensor of one (count, mean, M2) row per rank (12 C bytes) + 1 all_reduce (2 C floats, backward); per step "
                       "1 all_reduce of the parameter gradients (0.49 M floats), 3 scalar all_reduces (losses), 2 all_reduces of 3 floats (bounding box)",
    }


def run_partition(args, rank, world, local_rank):
    """``--mode partition``: the partitioned run alone, as the bench line."""
    from semigcn_b200.networks import SGCN_WIDTHS
    r = partition_measure(args, rank, world, local_rank, args.freq, args.steps, args.warmup)
    if r is None:
        return None
    return {
        "metric": METRIC, "value": r["edges_per_s"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"SGCN ({args.conv}, 13 blocks, widths {SGCN_WIDTHS}) fwd + step losses (position + face normals across cuts) + bwd + "
                               f"gradient all-reduce + Adam on ONE icosphere n={args.freq} ({r['vertices']} vertices, {r['directed_edges']} directed edges) "
                               f"vertex-partitioned over {world} GPU(s); BASELINE.json configs[3]",
                   "vertices": r["vertices"], "directed_edges": r["directed_edges"], "conv_layers": N_LAYERS,
                   "parallelism": f"vertex partition x{world}: " + r["collectives"],
                   "halo_rows_total": r["halo_rows_total"], "halo_rows_max_per_rank": r["halo_rows_max_per_rank"], "vertex_order": r["vertex_order"],
                   "overlap": r["overlap"], "l2_policy": "inputs larger than L2; no explicit flush"},
        "train_steps_per_s": 1e3 / r["ms_per_step"],
        "e2e": {"value": r["e2e_edges_per_s"], "unit": UNIT, "h2d_bytes_per_step": r["h2d_bytes_per_step"], "d2h_bytes_per_step": 8 * world,
                "ms_per_step": r["e2e_ms_per_step"]},
        "gpu_launches": r["gpu_launches"], "clocks": r["clocks"], "final_loss": r["final_loss"], "partition": r,
    }


def partition_block(args, rank, world, local_rank):
    """The partitioned runs carried by the default ``--gpus N`` line (N > 1):
      * the 3 994 242-vertex mesh (the largest of the series that also fits ONE GPU) partitioned over the N GPUs, against the
        unpartitioned step on the same mesh measured by rank 0 alone in this invocation  -> like-for-like strong scaling;
      * N >= 4: the 16 002 252-vertex mesh (configs[3]; needs >= 2 GPUs: ~260 GB of saved activations), against N x the
        single-GPU rate on a mesh of the PER-RANK size (16 M / N vertices) and against the 4 M single-GPU rate."""
    dev = torch.device(f"cuda:{local_rank}")
    steps, warmup = min(args.steps, 5), 3
    out = {}
    single = None
    if rank == 0:
        single = single_gpu_measure(args, dev, 632, steps, warmup)
    torch.distributed.barrier()
    p4 = partition_measure(args, rank, world, local_rank, 632, steps, warmup, e2e=False)
    if rank == 0:
        p4["single_gpu_same_mesh"] = single
        p4["speedup_vs_single_gpu_same_mesh"] = single["ms_per_step"] / p4["ms_per_step"]
        out["mesh_4m"] = p4
    if world >= 4:
        per_rank = None
        if rank == 0:
            import math
            f_pr = int(round(math.sqrt((16002252 / world - 2) / 10.0)))
            per_rank = single_gpu_measure(args, dev, f_pr, steps, warmup)
        torch.distributed.barrier()
        p16 = partition_measure(args, rank, world, local_rank, 1265, steps, warmup, e2e=False)
        if rank == 0:
            p16["single_gpu_per_rank_size_mesh"] = per_rank
            p16["speedup_vs_single_gpu_rate_at_per_rank_size"] = p16["edges_per_s"] / per_rank["edges_per_s"]
            p16["speedup_vs_single_gpu_rate_4m_mesh"] = p16["edges_per_s"] / single["edges_per_s"]
            p16["target"] = ">= 6x at 8 GPUs (BASELINE.json north_star)"
            out["mesh_16m"] = p16
    return out if rank == 0 else None


# ------------------------------------------------------------------------------------------
# extra blocks of the default line: BASELINE.json configs[0], [1], [4]
# ------------------------------------------------------------------------------------------
def graph_step_block(args, dev, reps: int = 10):
    """The headline step (same mesh, network, losses, optimizer) replayed as ONE CUDA graph (semigcn_b200/graphed.py): the narrow
    encoder / decoder layers of the eager step are launch-bound (20-80 us kernels behind ~60 us of Python per launch); the replay
    shows the step without those gaps.  Inputs (mask column) are copied into the graph's static buffers every step."""
    from semigcn_b200.graphed import GraphedTrainStep
    from semigcn_b200.networks import SingleScaleGCN
    prob = make_problem(args.freq, dev, seed=314, order=args.mesh_order)
    mesh = prob["mesh"]
    torch.manual_seed(314)
    net = SingleScaleGCN(dev, conv=args.conv).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=0.01, capturable=True)
    g = GraphedTrainStep(net, lambda o: step_losses(o, prob), opt, prob["z1"], prob["x_pos"], mesh.edge_index, prob["dms"][:, 0:1].contiguous())
    ms, _ = _time_steps(lambda i: g(prob["dms"][:, i % 8:i % 8 + 1]), reps, 3, torch.cuda.synchronize)
    return {"vertices": mesh.num_vertices, "directed_edges": mesh.nnz, "ms_per_step": ms / reps, "train_steps_per_s": 1e3 * reps / ms,
            "edges_per_s": float(mesh.nnz) * N_LAYERS * reps / (ms / 1e3), "kernels_per_step": g.launches_per_replay}


def small_mesh_block(args, dev, freq: int = 32, reps: int = 20):
    """configs[0] / configs[1] sizes (the reference's own meshes: 10-30 k vertices), one GPU: the SGCN step with the live
    conv of the reference (ChebConv, util/networks.py:13) and with GCNConv, and the MGCN step (util/meshnet.py mirror on a
    synthetic 3-level hierarchy), eager and replayed as ONE CUDA graph (the step is launch-bound at this size)."""
    from semigcn_b200 import losses, meshgen, ops
    from semigcn_b200.data import Data
    from semigcn_b200.graphed import GraphedTrainStep
    from semigcn_b200.meshnet import MGCN
    from semigcn_b200.networks import SingleScaleGCN
    from semigcn_b200.nn import MeshPool
    out = {}

    def timeit(fn):
        ms, _ = _time_steps(fn, reps, 3, torch.cuda.synchronize)
        return ms / reps

    prob = make_problem(freq, dev)
    mesh = prob["mesh"]
    dm0 = prob["dms"][:, 0:1].contiguous()
    for conv in ("chebconv", "gcnconv"):
        torch.manual_seed(314)
        net = SingleScaleGCN(dev, conv=conv).to(dev)
        opt = torch.optim.Adam(net.parameters(), lr=0.01, capturable=True)
        data = Data(z1=prob["z1"], x_pos=prob["x_pos"], edge_index=mesh.edge_index)

        def eager(i):
            opt.zero_grad(set_to_none=True)
            loss = step_losses(net(data, prob["dms"][:, i % 8:i % 8 + 1]), prob)
            loss.backward()
            opt.step()
            return loss

        ms_eager = timeit(eager)
        g = GraphedTrainStep(net, lambda o: step_losses(o, prob), opt, prob["z1"], prob["x_pos"], mesh.edge_index, dm0)
        ms_graph = timeit(lambda i: g(prob["dms"][:, i % 8:i % 8 + 1]))
        out[f"sgcn_{conv}"] = {"vertices": mesh.num_vertices, "directed_edges": mesh.nnz, "eager_ms_per_step": ms_eager, "graph_ms_per_step": ms_graph,
                               "train_steps_per_s": 1e3 / ms_graph, "edges_per_s": float(mesh.nnz) * N_LAYERS / (ms_graph / 1e3),
                               "kernels_per_step": g.launches_per_replay}
        del g, net, opt
    # MGCN: 4 resolutions (N, 0.6 N, 0.36 N, 0.216 N), 33 ChebConv(K=3) layers
    sp = meshgen.synth_inpainting_problem(freq, device=dev, smooth_iters=10, n_dummy=8)
    smesh = sp["mesh"]
    hier = meshgen.synth_pool_hierarchy(smesh)
    sm, ini, vmask = [sp["x_pos"]], [sp["ini_vs"].float()], [sp["v_mask"]]
    for lvl in range(3):
        pool = MeshPool(hier["p_hashes"][lvl]).to(dev)
        sm.append(pool(sm[-1]))
        ini.append(pool(ini[-1]))
        vmask.append(pool(vmask[-1].float().reshape(-1, 1)).reshape(-1) == 1.0)
    torch.manual_seed(314)
    net = MGCN(dev, hier["edge_inds"], hier["p_hashes"], hier["up_hashes"], sm, skip=False, drop_rate=0.0, tensor_masks=True).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=0.01, capturable=True)
    pos_w = [0.35, 0.3, 0.2, 0.15]                                                     # mgcn.py:82

    def mgcn_loss(poss):
        lp = sum(losses.fused_mask_pos_rec_loss(p, t, m) * w for p, t, m, w in zip(poss, ini, vmask, pos_w))
        _, ln = losses.sgcn_step_losses(poss[0], smesh.faces, sp["ini_vs"], sp["fn"], sp["v_mask"], sp["f_mask"])
        return lp + K1 * ln

    data = Data(z1=sp["z1"], x_pos=sp["x_pos"])
    dm = sp["vmask_dummy"][:, :1].contiguous()

    def mgcn_eager(i):
        opt.zero_grad(set_to_none=True)
        loss = mgcn_loss(net(data, dm))
        loss.backward()
        opt.step()
        return loss

    ms_eager = timeit(mgcn_eager)
    g = GraphedTrainStep(net, mgcn_loss, opt, sp["z1"], sp["x_pos"], None, dm)
    ms_graph = timeit(lambda i: g(dm))
    nnz_l = [int(e.shape[1]) for e in hier["edge_inds"]]
    out["mgcn_chebconv"] = {"vertices_per_level": hier["sizes"], "directed_edges_per_level": nnz_l, "conv_layers": 33,
                            "eager_ms_per_step": ms_eager, "graph_ms_per_step": ms_graph, "train_steps_per_s": 1e3 / ms_graph,
                            "kernels_per_step": g.launches_per_replay}
    del g, net, opt
    ops.clear_graph_cache()
    torch.cuda.empty_cache()
    return out


def mesh_order_block(args, dev):
    """the headline runs on the generator's vertex numbering (subdivision order, row by row inside each
    icosahedron face); this is the same step on the SAME mesh renumbered along a Morton (Z-order) curve -- compact patches
    instead of mesh rows per CTA, i.e. more neighbour rows served by L1 in the aggregation kernel.  A dataset property, not a
    kernel change: reported next to the headline, never instead of it."""
    from semigcn_b200 import ops, profile
    from semigcn_b200.data import Data
    from semigcn_b200.networks import SingleScaleGCN
    prob = make_problem(args.freq, dev, seed=314, order="morton")
    mesh = prob["mesh"]
    torch.manual_seed(314)
    net = SingleScaleGCN(dev, conv=args.conv).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=0.01)
    data = Data(z1=prob["z1"], x_pos=prob["x_pos"], edge_index=mesh.edge_index)

    def one_step(i):
        opt.zero_grad(set_to_none=True)
        loss = step_losses(net(data, prob["dms"][:, i % 8:i % 8 + 1]), prob)
        loss.backward()
        opt.step()
        return loss

    steps = min(args.steps, 10)
    ms, _ = _time_steps(one_step, steps, 3, torch.cuda.synchronize)
    with profile.KernelProfile() as kp:
        for i in range(3):
            one_step(i)
        fam = kp.summary()
    spmm_ms = sum(v["ms"] for k, v in fam.items() if k.startswith("spmm_c")) / 3
    spmm_b = sum(v["bytes"] for k, v in fam.items() if k.startswith("spmm_c")) / 3
    out = {"vertices": mesh.num_vertices, "ms_per_step": ms / steps, "edges_per_s": float(mesh.nnz) * N_LAYERS * steps / (ms / 1e3),
           "spmm_ms_per_step": spmm_ms, "spmm_GBps": spmm_b / (spmm_ms / 1e3) / 1e9 if spmm_ms else None}
    del net, opt, data, prob
    ops.clear_graph_cache()
    torch.cuda.empty_cache()
    return out


def batch64_block(args, rank, world, local_rank, meshes_total: int = 64, freq: int = 100, steps_per_mesh: int = 10, slots: int = 4):
    """configs[4]: a batch of 64 independent 100 k-vertex meshes, one self-prior network per mesh (sgcn.py:78-80), spread over
    the N GPUs: rank r trains meshes r, r + N, ...  No data-path collective.  On each GPU ``slots`` meshes train CONCURRENTLY,
    one CUDA stream each, every step a whole-step CUDA graph (the 100 k-vertex step is launch-bound: DESIGN.md §5.7); when
    a mesh is done the slot's graph is reused for the next one -- new inputs / targets copied into the static buffers,
    parameters and optimizer state reset in place.  ``steps_per_mesh`` steps per mesh are timed (a bounded sample of the
    4 000 steps of a real run).  Aggregate over ranks, max-over-ranks time."""
    from semigcn_b200 import ops
    from semigcn_b200.graphed import GraphedTrainStep
    from semigcn_b200.networks import SingleScaleGCN
    dev = torch.device(f"cuda:{local_rank}")
    mine = list(range(rank, meshes_total, world))
    slots = max(1, min(slots, len(mine)))
    lanes = []
    for sidx in range(slots):
        prob = make_problem(freq, dev, seed=1000 + mine[sidx])
        torch.manual_seed(314 + mine[sidx])
        net = SingleScaleGCN(dev, conv=args.conv).to(dev)
        opt = torch.optim.Adam(net.parameters(), lr=0.01, capturable=True)
        init = [p.detach().clone() for p in net.parameters()] + [b.detach().clone() for b in net.buffers()]
        g = GraphedTrainStep(net, (lambda pr: (lambda o: step_losses(o, pr)))(prob), opt, prob["z1"], prob["x_pos"], prob["mesh"].edge_index,
                             prob["dms"][:, 0:1].contiguous())
        lanes.append(dict(prob=prob, net=net, opt=opt, init=init, graph=g, stream=torch.cuda.Stream(device=dev), queue=mine[sidx::slots]))
    nnz = lanes[0]["prob"]["mesh"].nnz
    nv = lanes[0]["prob"]["mesh"].num_vertices
    # per-mesh inputs: same topology (icosphere n=100), different geometry noise -> generated up front, copied in when a mesh starts
    gen = torch.Generator(device=dev)

    def mesh_inputs(mesh_id, prob):
        gen.manual_seed(1000 + mesh_id)
        bump = torch.randn(nv, 1, generator=gen, dtype=torch.float64, device=dev)
        return prob["mesh"].vs * (1.0 + 0.02 * bump)

    def start_mesh(lane, mesh_id):
        from semigcn_b200 import meshgen
        pr = lane["prob"]
        with torch.no_grad():
            ini = mesh_inputs(mesh_id, pr)
            smo = meshgen.uniform_laplacian_smooth(ini, pr["mesh"].edge_index, 30)
            pr["ini"].copy_(ini)
            pr["fn"].copy_(meshgen.face_normals(ini, pr["mesh"].faces))
            lane["graph"].z1.copy_((ini - smo).float())
            lane["graph"].x_pos.copy_(smo.float())
            for t, v in zip(list(lane["net"].parameters()) + list(lane["net"].buffers()), lane["init"]):
                t.copy_(v)
            for st in lane["opt"].state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()

    def run_all():
        cur = torch.cuda.current_stream(dev)
        for lane in lanes:
            lane["stream"].wait_stream(cur)
        rounds = max(len(l["queue"]) for l in lanes)
        for rnd in range(rounds):
            for lane in lanes:
                if rnd < len(lane["queue"]):
                    with torch.cuda.stream(lane["stream"]):
                        start_mesh(lane, lane["queue"][rnd])
            for i in range(steps_per_mesh):
                for lane in lanes:
                    if rnd < len(lane["queue"]):
                        with torch.cuda.stream(lane["stream"]):
                            lane["graph"](lane["prob"]["dms"][:, i % 8:i % 8 + 1])
        for lane in lanes:
            cur.wait_stream(lane["stream"])

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for lane in lanes:                       # warm-up: one mesh start + two steps per lane
        with torch.cuda.stream(lane["stream"]):
            start_mesh(lane, lane["queue"][0])
            lane["graph"](lane["prob"]["dms"][:, 0:1])
            lane["graph"](lane["prob"]["dms"][:, 1:2])
    barrier()
    sampler.mark_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_all()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = sum(l["graph"].launches_per_replay for l in lanes[:1]) * steps_per_mesh * len(mine)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t[0])
    del lanes
    ops.clear_graph_cache()
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    total_steps = meshes_total * steps_per_mesh
    return {"meshes": meshes_total, "vertices_per_mesh": nv, "directed_edges_per_mesh": nnz, "n_gpus": world, "meshes_per_gpu": len(mine),
            "concurrent_meshes_per_gpu": slots, "steps_per_mesh_timed": steps_per_mesh, "ms_total": ms, "train_steps_per_s": total_steps / (ms / 1e3),
            "edges_per_s": float(nnz) * N_LAYERS * total_steps / (ms / 1e3), "ms_per_step_per_gpu": ms / (len(mine) * steps_per_mesh),
            "gpu_launches_rank0": launches, "clocks": clocks, "collectives": "none (independent meshes)",
            "includes": "per mesh: geometry + 30 smoothing iterations + target normals on the GPU, in-place reset of parameters / Adam state, "
                        f"then {steps_per_mesh} whole-step CUDA-graph replays (forward + both step losses + backward + Adam)"}


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
    if args.mode == "replicas" and not args.no_extras and not args.cuda_graph:
        # extra blocks (every rank takes part; rank 0 holds the results).  They must never cost the bench line.
        from semigcn_b200 import ops
        ops.clear_graph_cache()
        torch.cuda.empty_cache()
        extras = {}

        def guarded(name, fn):
            try:
                r = fn()
                if r is not None:
                    extras[name] = r
            except Exception as exc:   # noqa: BLE001
                extras[name] = {"error": repr(exc)[:400]}
                ops.clear_graph_cache()
                torch.cuda.empty_cache()

        if world > 1:
            guarded("partition", lambda: partition_block(args, rank, world, local_rank))
        guarded("batch64", lambda: batch64_block(args, rank, world, local_rank))
        if world == 1:
            guarded("cuda_graph_step", lambda: graph_step_block(args, torch.device(f"cuda:{local_rank}")))
            ops.clear_graph_cache()
            torch.cuda.empty_cache()
            guarded("small_meshes", lambda: small_mesh_block(args, torch.device(f"cuda:{local_rank}")))
            guarded("mesh_order_morton", lambda: mesh_order_block(args, torch.device(f"cuda:{local_rank}")))
        if line is not None:
            line.update(extras)
    if line is not None:
        print(json.dumps(line), flush=True)
    if dist_on:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
