#!/usr/bin/env python
"""bench.py -- STINet hot path (multi-level mesh U-Net forward + backward over batched mesh graphs) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl stinet|reference] [--workload cfg2] [--dtype fp32]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

metric  = level-0 mesh vertices per second for one training step (graph-structure build + forward + masked-L1 loss +
          backward [+ gradient all-reduce when N>1] + Adam step), BASELINE.json `metric`, on BASELINE.json configs[1]:
          8 x 40,962-vertex synthetic icosphere crops, 10 input channels, 4 trace-map levels, fp32.
value   = whole-job throughput, inputs resident in HBM when the timed region starts (weak scaling: every rank owns
          its own batch, value = total vertices / max-over-ranks device time).
e2e     = same step through the public module API with HOST (pinned) inputs: H2D of the batch and a D2H read of the
          loss inside the timed region.
roofline= the dominant kernel of the step (largest share of device time in a CUDA-event profile of the same step),
          algorithmic bytes or flops per launch / its mean CUDA-event duration, against MEASURED_PEAKS.json.
cpu_baseline / --impl reference = the CPU oracle port (oracle/stinet_oracle.py; PyG is not installable, so the
          reference itself cannot run) on the host cores, on a bounded sample (ONE graph of the batch).
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys
import threading
import time

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1 when the
# box sets NCCL_DEBUG), so main() sets the real stdout aside for the result line and points fd 1 at stderr for
# everything else.
_RESULT_OUT = None


def reserve_stdout() -> None:
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(result: dict) -> None:
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(result) + "\n")
    out.flush()


ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "surface-texture-inpainting-net_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "cfg2": dict(kind="icosphere", gen=dict(subdiv=6), batch=8,
                 net=dict(input_nc=10, output_nc=3, ngf=64, filter_type="edgeconvtransinv", norm="instance", n_blocks=9,
                          n_levels=4, pooling_type="max", checkpoint_bottleneck=True),
                 name="STINet 3D, synthetic icosphere crops 8 x 40,962 vertices, 10 ch, 4 trace-map levels, ngf 64, 9 blocks"),
    # configs[0]: the reference's CPU-runnable 2D case (parity-test case, selectable for local runs)
    "cfg1": dict(kind="grid", gen=dict(size=128), batch=4,
                 net=dict(input_nc=4, output_nc=3, ngf=64, filter_type="edgeconv", norm="instance", n_blocks=9,
                          n_levels=4, pooling_type="max"),
                 name="STINet 2D, 4 x 128x128 image-grid graphs, 4 pool levels"),
    # configs[2]: ScanNet-scale scene, full-scene inference (eval / no_grad, batch=None norms)
    "cfg3": dict(kind="plane", gen=dict(rows=500, cols=500, mask_cover=0.05), batch=1, mode="infer",
                 net=dict(input_nc=10, output_nc=3, ngf=64, filter_type="edgeconvtransinv", norm="instance", n_blocks=9,
                          n_levels=4, pooling_type="max"),
                 name="STINet 3D full-scene inference, synthetic 250,000-vertex plane mesh (1.5 M edges), 4 trace-map levels"),
    # configs[4]: one ~2 M-vertex scene per GPU, 5 trace-map levels (run with --dtype bf16)
    "cfg5": dict(kind="plane", gen=dict(rows=1448, cols=1448, mask_cover=0.0001, mask_radius=8), batch=1,
                 net=dict(input_nc=10, output_nc=3, ngf=64, filter_type="edgeconvtransinv", norm="instance", n_blocks=9,
                          n_levels=5, pooling_type="max"),
                 name="STINet 3D, one synthetic 2,096,704-vertex plane mesh (12.6 M edges) per GPU, 5 trace-map levels"),
    "tiny": dict(kind="icosphere", gen=dict(subdiv=3, mask_radius=3), batch=2,
                 net=dict(input_nc=10, output_nc=3, ngf=16, filter_type="edgeconvtransinv", norm="instance", n_blocks=2,
                          n_levels=2, pooling_type="max"),
                 name="tiny debug workload"),
}
# tcgen05 passes per algorithmic product and the tensor-pipe rate of the operand type relative to bf16 / fp16
MODE_COST = {"fp32": (3, 1.0), "f16": (1, 1.0), "bf16x3": (3, 1.0), "bf16": (1, 1.0), "tf32": (1, 0.5), "fp32_tf32x3": (3, 0.5)}
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return dict(FALLBACK_PEAKS), "fallback"


def path_roofline(summ: dict, steps: int, hbm_gbs: float, bf16_tflops: float, passes: int, rate: float) -> dict:
    """the path's roofline time = sum over kernels of max(algorithmic bytes / HBM peak, flops / tensor peak).
    `summ` = KernelProfiler.summary() of `steps` eager steps.  Two tensor ceilings are reported: the dense bf16 peak,
    and the ceiling of the arithmetic mode in use (peak * rate / passes, e.g. 3 TF32 passes at half rate for fp32)."""
    t_bf16 = t_mode = t_meas = 0.0
    for key, r in summ.items():
        t_b = r["bytes"] / (hbm_gbs * 1e9)
        dense = key.startswith("linear")
        t_f = r["flops"] / (bf16_tflops * 1e12) if dense else 0.0
        t_bf16 += max(t_b, t_f)
        t_mode += max(t_b, t_f * passes / rate)
        t_meas += r["ms"] * 1e-3
    per = 1e3 / max(steps, 1)
    return {"roofline_ms_per_step": t_bf16 * per, "roofline_ms_per_step_mode_ceiling": t_mode * per,
            "measured_ms_per_step_eager": t_meas * per,
            "note": "sum over kernels of max(algorithmic bytes / HBM peak, flops / tensor peak); HBM bytes follow the "
                    "no-cache-reuse convention of SURVEY 8d, so this is an upper bound on the time a perfect "
                    "implementation needs"}


def masked_l1(out, b):
    """the trainer's loss, reference trainers/inpainting3d_trainer.py:127-137 (torch.where + L1 + 0.99^mask + mean), as the
    package's fused op (one forward and one backward kernel; tests/test_gpu_kernels.py checks it against the torch chain)"""
    from stinet_b200 import ops
    return ops.masked_l1_loss(out, b.color, b.mask)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region through NVML in a background thread (the same
    counters as the recipe's nvidia-smi line in B200_PROFILING.md, without forking nvidia-smi every 100 ms, which
    stalls kernel launches)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period, self.samples, self.err = index, period_s, [], None
        self._stop = threading.Event()
        self.thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((mhz, rs, pw))
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            self._stop.wait(self.period)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"]}
        self._stop.set()
        self.thread.join(timeout=2)
        sm = sorted(m for m, _, _ in self.samples)
        reasons = sorted(n for n, bit in self.REASONS.items() if any(r & bit for _, r, _ in self.samples))
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                "power_w_max": max((p for _, _, p in self.samples), default=None), "reasons": reasons}


# ----------------------------------------------------------------------------------------------------------------


REFERENCE_BUDGET_S = 300.0      # wall-clock bound of one --impl reference run (the driver calls it at every N)


def _oracle_job(wl, n_graphs):
    """(net, batch, step) of the CPU oracle port on `n_graphs` graphs of the workload: fwd + masked-L1 + bwd (recompute
    off), or the forward alone for the inference workload."""
    from oracle import stinet_oracle as O
    from stinet_b200 import synthetic
    torch.manual_seed(49)
    nk = {("norm_type" if k == "norm" else k): v for k, v in wl["net"].items()}
    net = O.OracleSTINet(**nk)
    b = synthetic.make_batch(wl["kind"], n_graphs, wl["net"]["n_levels"], seed=49, **wl["gen"])
    infer = wl.get("mode") == "infer"

    def step():
        if infer:
            with torch.no_grad():
                net(b)
            return
        net.zero_grad(set_to_none=True)
        O.masked_l1_loss(net(b), b).backward()

    return net, b, step


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's algorithm for the path on the host cores, all threads.  torch_geometric /
    torch_scatter cannot be installed in this image, so this is the CPU oracle port (kind='port').  A step is the WHOLE
    batch of the workload (same config as the stinet arm: per-graph norms, B graphs); --steps / --warmup are honoured and
    only cut short if the run would exceed REFERENCE_BUDGET_S (then the line says how many steps were timed)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    infer = wl.get("mode") == "infer"
    _, b, step = _oracle_job(wl, wl["batch"])
    n0 = b.x.shape[0]
    t_begin = time.perf_counter()
    step()                                                   # first warm-up step, also calibrates the budget
    t_one = time.perf_counter() - t_begin
    warm = max(1, args.warmup)
    steps = max(1, args.steps)
    fit = int(REFERENCE_BUDGET_S / max(t_one, 1e-3))         # steps (warm-up included) that fit the budget
    if warm + steps > fit:
        warm = max(1, min(warm, fit // 4))
        steps = max(1, fit - warm)
    for _ in range(warm - 1):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    v = n0 / dt
    what = "forward (eval, no_grad)" if infer else "fwd+loss+bwd"
    sample = (f"whole batch: {wl['batch']} graph(s), {n0} vertices, {what}, {steps} timed steps after {warm} warm-up"
              + ("" if (steps == max(1, args.steps) and warm == max(1, args.warmup)) else
                 f" (asked for {args.steps}/{args.warmup}; cut to fit {REFERENCE_BUDGET_S:.0f} s)"))
    emit({
        "impl": "reference", "metric": "mesh vertices/sec fwd" if infer else "mesh vertices/sec fwd+bwd", "value": v, "unit": "vertices/s", "n_gpus": 0,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "vertices_per_step_per_gpu": n0, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "vertices/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "vertices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def cpu_baseline(wl, budget_s=25.0):
    """The same CPU oracle port on a bounded sample for the stinet arm's JSON line: the whole batch when one step of it
    fits the budget twice, else one graph."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    infer = wl.get("mode") == "infer"
    _, b1, step1 = _oracle_job(wl, 1)
    t0 = time.perf_counter()
    step1()                                      # warm-up on ONE graph, also calibrates the budget
    t_graph = time.perf_counter() - t0
    n_graphs = wl["batch"] if 3 * t_graph * wl["batch"] <= budget_s else 1
    if n_graphs == 1:
        b, step = b1, step1
    else:
        _, b, step = _oracle_job(wl, n_graphs)
        step()
    n = max(1, min(5, int(budget_s / max(t_graph * n_graphs, 1e-3)) - 1))
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    dt = (time.perf_counter() - t0) / n
    return {"value": b.x.shape[0] / dt, "unit": "vertices/s", "cores": cores, "kind": "port",
            "sample": f"CPU oracle port (PyG absent), {n_graphs} of {wl['batch']} graphs ({b.x.shape[0]} vertices), "
                      f"{'forward (eval, no_grad)' if infer else 'fwd+loss+bwd'}, "
                      f"{n} timed steps after 1 warm-up, fp32, recompute off"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="stinet", choices=["stinet", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "f16", "bf16", "bf16x3", "bf16_1pass", "tf32", "tf32x3"],
                    help="arithmetic of the dense layers (tcgen05, fp32 accumulate): fp32 = three kind::f16 passes on scaled fp16 "
                         "hi/lo operand planes (fp32-class, the 1e-5 parity mode); f16 (alias bf16: the 16-bit mode BASELINE's "
                         "config 5 asks for) = ONE kind::f16 pass on the hi planes (11-bit operands: outputs within 2e-2); "
                         "bf16x3 / bf16_1pass = bf16 tiles (3 passes / 1 pass); tf32x3 = round 1's 3xTF32 mode; tf32 = one pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (used under ncu only)")
    ap.add_argument("--no-cached", action="store_true", help="skip the cached-structure leg")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch every kernel from Python instead of replaying the captured whole-step CUDA graph")
    ap.add_argument("--kernels-out", default=None, help="write the full per-entry-point CUDA-event table (JSON) here")
    ap.add_argument("--profiler-range", action="store_true",
                    help="bracket timed region 1 with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    args = ap.parse_args()
    reserve_stdout()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    from stinet_b200 import _abi, synthetic
    from stinet_b200.models import surfacetextureinpaintingnet as S
    from stinet_b200.parallel import GradAllReducer, init_distributed

    rank, local, world = init_distributed()
    assert torch.cuda.is_available(), "bench.py --impl stinet needs a CUDA device (there is no CPU fallback)"
    assert _abi.load().stinet_device_ok() == 1, _abi.load().stinet_last_error().decode()
    dev = torch.device("cuda", local)
    W, K = max(args.warmup, 3), args.steps

    torch.manual_seed(49)                                    # same initial weights on every rank
    precision = {"fp32": "fp32", "f16": "f16", "bf16": "f16", "bf16x3": "bf16x3", "bf16_1pass": "bf16", "tf32": "tf32",
                 "tf32x3": "fp32_tf32x3"}[args.dtype]
    infer = wl.get("mode") == "infer"
    net = S.define_G(**wl["net"], gpu_ids=[dev], precision=precision)
    net = net.eval() if infer else net.train()
    host = synthetic.make_batch(wl["kind"], wl["batch"], wl["net"]["n_levels"], seed=49 + 1000 * rank, **wl["gen"])
    host = host.pin_memory()
    n0 = int(host.x.shape[0])
    resident = host.to(dev)
    reducer = GradAllReducer(net)
    opt = torch.optim.Adam(net.parameters(), lr=7e-5, amsgrad=True, fused=True, capturable=True)  # shipped 3D config :95-101

    def eager_step(b, read_loss=False):
        b = copy.copy(b)
        b.__dict__.pop("_stinet_cache", None)                # every step rebuilds the graph structure (new batch)
        if infer:                                            # full-scene inference: structure build + forward
            with torch.no_grad():
                out = net(b)
            return float(out[0, 0].item()) if read_loss else out
        reducer.zero_grad()
        loss = masked_l1(net(b), b)
        (loss * reducer.loss_scale if world > 1 else loss).backward()
        reducer.finish()
        opt.step()
        return loss.item() if read_loss else loss

    graphed = None
    if args.no_graph:
        step = eager_step
    elif infer:
        # full-scene inference through the public graphed-forward API: structure build + forward captured per batch shape
        from stinet_b200.engine import GraphedForward
        graphed = GraphedForward(net)

        def step(b, read_loss=False):
            out = graphed(b)
            return float(out[0, 0].item()) if read_loss else out
    else:
        # the public whole-step API: the same work (structure build + fwd + loss + bwd + Adam) captured once per batch
        # shape into a CUDA graph and replayed; inputs are copied into the graph's static buffers on every call
        from stinet_b200.engine import GraphedTrainStep
        graphed = GraphedTrainStep(net, masked_l1, opt, reducer)

        def step(b, read_loss=False):
            loss = graphed(b)
            return loss.item() if read_loss else loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(W):
        step(resident)
    # ---- timed region 1: inputs resident in HBM -------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    def launch_count():
        return _abi.query("stinet_launch_count") + (graphed.replayed_launches if graphed is not None else 0)

    launches0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.profiler_range:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(K):
        step(resident)
    e1.record()
    barrier()
    if args.profiler_range:
        torch.cuda.profiler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    value = world * n0 * K / (ms * 1e-3)

    # ---- timed region 2: end to end through the public API from pinned host memory ----------------------------
    e2e = None
    if not args.no_e2e:
        if graphed is None or infer:
            def e2e_step():                                  # blocking H2D of the pinned batch, result element read back
                return step(host.to(dev, non_blocking=True) if graphed is None else host, read_loss=True)
        else:
            # the loader pattern: the pinned batch of step k+1 is copied host->device on a copy stream while step k
            # runs (GraphedTrainStep.prefetch); every step still moves its own inputs H2D and reads its loss D2H
            def e2e_step():
                loss = graphed(host)
                graphed.prefetch(host)
                return loss.item()
            graphed.prefetch(host)
        for _ in range(2):
            e2e_step()
        barrier()
        e0.record()
        for _ in range(K):
            e2e_step()
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        e2e = {"value": world * n0 * K / (ms_e2e * 1e-3), "unit": "vertices/s",
               "h2d_bytes_per_step": host.tensor_bytes(), "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / K,
               "pipeline": "python launches, blocking H2D" if graphed is None else
                           "H2D of the pinned batch into the graph's static inputs, graph replay, one output element read back" if infer else
                           "H2D of batch k+1 (pinned) overlaps step k on a copy stream; loss read back every step"}

    # ---- timed region 3 (extra key, not the headline): structure cached per sample ----------------
    # CSR / cluster CSR of every sample built once (dataset set-up, untimed); per step the batch structure is their
    # block-diagonal concatenation on the device, the host sends features only (no int64 COO tensors), and the captured
    # graph holds no structure-build kernels.  Same loader pattern and same H2D / D2H accounting as `e2e`.
    cached = None
    if graphed is not None and not infer and not args.no_e2e and not args.no_cached:
        from stinet_b200.data import collate
        from stinet_b200.structure import SampleStructure
        samples = synthetic.make_samples(wl["kind"], wl["batch"], wl["net"]["n_levels"], seed=49 + 1000 * rank, **wl["gen"])
        structs = [SampleStructure.build(s_, wl["net"]["n_levels"], dev) for s_ in samples]
        lean = collate(samples, keep_index=False).pin_memory()

        def fresh():
            return copy.copy(lean)                           # a new batch object per step, as a loader yields

        nxt = fresh()
        graphed.prefetch(nxt, structs)                       # attaches the structure; shape not captured yet
        graphed(nxt)                                         # capture
        nxt = fresh()
        graphed.prefetch(nxt, structs)

        def cached_step():
            nonlocal nxt
            loss = graphed(nxt)
            nxt = fresh()
            graphed.prefetch(nxt, structs)
            return loss.item()

        for _ in range(2):
            cached_step()
        barrier()
        e0.record()
        for _ in range(K):
            cached_step()
        e1.record()
        barrier()
        ms_c = max_over_ranks(e0.elapsed_time(e1))
        cached = {"value": world * n0 * K / (ms_c * 1e-3), "unit": "vertices/s", "ms_per_step": ms_c / K,
                  "h2d_bytes_per_step": lean.tensor_bytes(host_only=True), "d2h_bytes_per_step": 4,
                  "structure_bytes_in_hbm_per_batch": sum(s_.nbytes() for s_ in structs),
                  "what": "e2e with per-sample cached CSR: device-side block-diagonal concat per step, features-only H2D"}

    # ---- roofline leg: CUDA-event profile of the same step (outside the timed regions) ----------------------------
    pk, pk_src = peaks()
    roofline, kernels = None, None
    if world > 1 and rank != 0 and not args.no_profile:
        for _ in range(2):                                   # the eager steps all-reduce: every rank has to take part
            eager_step(resident)
    if rank == 0 and not args.no_profile:
        # a kernel's duration is taken with the kernel alone on the GPU: the profiled eager steps use the one-stream
        # schedule (the timed step overlaps the structure build and the weight gradients with the main chain, which
        # stretches every event pair that brackets two kernels sharing the SMs)
        from stinet_b200 import ops as _ops
        side_env, side_wgrad = os.environ.get("STINET_STRUCT_SIDE_STREAM"), _ops._WGRAD_SIDE
        os.environ["STINET_STRUCT_SIDE_STREAM"], _ops._WGRAD_SIDE = "0", 0
        try:
            with _abi.KernelProfiler() as prof:              # per-kernel CUDA events need eager launches
                for _ in range(2):
                    # let the host run ahead of the device: ~60 ms of idle spinning on the stream while Python enqueues
                    # the step's ~650 launches, so that the event pairs bracket device time, not the host's launch gaps
                    torch.cuda._sleep(120_000_000)
                    eager_step(resident)
        finally:
            _ops._WGRAD_SIDE = side_wgrad
            if side_env is None:
                os.environ.pop("STINET_STRUCT_SIDE_STREAM", None)
            else:
                os.environ["STINET_STRUCT_SIDE_STREAM"] = side_env
        summ = prof.summary()
        total = sum(r["ms"] for r in summ.values())
        kernels = {k: {"calls_per_step": r["calls"] // 2, "ms_per_step": round(r["ms"] / 2, 4),
                       "share": round(r["ms"] / total, 4),
                       "GBps": round(r["bytes"] / (r["ms"] * 1e-3) / 1e9, 1) if r["ms"] > 0 else None,
                       "TFLOPs": round(r["flops"] / (r["ms"] * 1e-3) / 1e12, 2) if r["ms"] > 0 else None}
                   for k, r in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])[:12]}
        if args.kernels_out:
            with open(args.kernels_out, "w") as f:
                json.dump({"total_ms_per_step": total / 2, "kernels": {
                    k: {"calls_per_step": r["calls"] // 2, "ms_per_step": r["ms"] / 2, "share": r["ms"] / total,
                        "GBps": r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["ms"] > 0 else None,
                        "TFLOPs": r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] > 0 else None}
                    for k, r in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])}}, f, indent=1)
        # dominant kernel = the kernel FAMILY (entry point, all shapes) with the largest share of device time
        # (the three dense-layer entry points fwd / dgrad / wgrad run ONE kernel, gemm_tc_kernel: one family)
        fam = {}
        for k, r in summ.items():
            name = k.split("[")[0]
            f = fam.setdefault("linear" if name.startswith("linear") else name, {"calls": 0, "ms": 0.0, "bytes": 0, "flops": 0})
            for q in f:
                f[q] += r[q]
        top, r = max(fam.items(), key=lambda kv: kv[1]["ms"])
        bf16_peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        if top.startswith("linear"):
            ach = r["flops"] / (r["ms"] * 1e-3) / 1e12
            # MMA passes per algorithmic product, and the tensor-pipe rate of the operand type relative to bf16
            passes, rate = MODE_COST[precision]
            roofline = {"kernel": "gemm_tc_kernel (dense layers: fwd + dgrad + wgrad, all shapes)", "bound": "tensor", "achieved": ach,
                        "peak": bf16_peak, "unit": "TFLOP/s", "frac": ach / bf16_peak, "traffic": None,
                        "peak_source": f"{pk_src} dense bf16 GEMM (sustained)",
                        "mode": f"{precision}: {passes} tcgen05 pass(es) per product at {rate}x the bf16 rate; achieved counts "
                                f"algorithmic flops 2MNK once, so the ceiling of this mode is peak*{rate / passes:.3f}",
                        "frac_of_mode_ceiling": ach / (bf16_peak * rate / passes)}
        else:
            ach = r["bytes"] / (r["ms"] * 1e-3) / 1e9
            roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": ach / pk["hbm_gbs"], "traffic": None, "peak_source": f"{pk_src} copy bandwidth"}
        try:                                               # never let a reporting extra take the bench down
            passes_, rate_ = MODE_COST[precision]
            roofline["path"] = path_roofline(summ, 2, pk["hbm_gbs"], bf16_peak, passes_, rate_)
            roofline["path"]["frac_of_timed_step"] = roofline["path"]["roofline_ms_per_step"] / (ms / K)
            roofline["path"]["frac_of_timed_step_mode_ceiling"] = \
                roofline["path"]["roofline_ms_per_step_mode_ceiling"] / (ms / K)
        except Exception as e:  # noqa: BLE001
            roofline["path"] = {"error": repr(e)}
        roofline["schedule"] = ("per-kernel durations: eager launches on one stream (each kernel alone on the GPU); the timed "
                                "step overlaps structure build and weight gradients with the main chain on side streams")
        roofline["share_of_step"] = r["ms"] / total
        roofline["ms_per_launch"] = r["ms"] / r["calls"]
        # measured DRAM traffic per launch of the dominant kernels (ncu --set full, profiles/): largest shape of each
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            ent = tj.get(args.workload, {}).get(top)
            if ent:
                roofline["traffic"] = ent.get("dram_bytes_per_launch")
                roofline["traffic_note"] = ent.get("note")
        # the HBM-bound families are always reported too (north_star: aggregation / pool / norm kernels vs HBM peak);
        # algorithmic bytes follow SURVEY 8d's no-cache-reuse convention, so L2 hits can push a fraction above 1
        hbm = {k: v for k, v in fam.items() if not k.startswith("linear") and v["ms"] > 0 and v["bytes"] > 0}
        roofline["hbm_kernels"] = {
            k: {"achieved_GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1),
                "frac": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / pk["hbm_gbs"], 3),
                "share_of_step": round(v["ms"] / total, 4), "launches_per_step": v["calls"] // 2}
            for k, v in sorted(hbm.items(), key=lambda kv: -kv[1]["ms"])}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(wl)

    if rank == 0:
        emit({
            "metric": "mesh vertices/sec fwd" if infer else "mesh vertices/sec fwd+bwd", "value": value,
            "unit": "vertices/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "f16": "f16", "bf16": "f16", "bf16x3": "bf16", "bf16_1pass": "bf16", "tf32": "tf32",
                      "tf32x3": "f32"}[args.dtype], "data": "synthetic",
            "config": {"workload": wl["name"], "vertices_per_step_per_gpu": n0,
                       "step": "graph-structure (CSR) build + forward (eval, no_grad)" if infer else
                               "graph-structure (CSR) build + forward + masked L1 + backward"
                               + (" + NCCL gradient all-reduce" if world > 1 else "") + " + Adam(amsgrad) step",
                       "dense_layers": precision,
                       "launch": "python launches" if graphed is None else
                                 ("forward CUDA graph replay (stinet_b200.engine.GraphedForward)" if infer else
                                  "whole-step CUDA graph replay (stinet_b200.engine)"),
                       "l2": "per-step working set (activations + weights, several GB) is far larger than the 126 MB L2; "
                             "no explicit flush"},
            "e2e": e2e, "e2e_cached_structure": cached, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
            "kernels": kernels,
        })
    if world > 1:
        # Leave without tearing NCCL down: the step graphs hold captured NCCL kernels, and destroy_process_group() with
        # such graphs alive can block for minutes (observed on 2 x B200).  Everything is measured and printed by now; all
        # ranks meet at a barrier, flush, and exit with status 0.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        if _RESULT_OUT is not None:
            _RESULT_OUT.flush()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
