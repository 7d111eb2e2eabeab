"""GPU check of the dense layers on fp16 operand planes (stinet_linear_{fwd,dgrad,wgrad}_f16, stinet_f16_{amax,split})
against an fp64 torch reference, next to the 3xTF32 entry points they replace: max-norm relative error, CUDA-event time
of the GEMM alone (planes ready) and of the amax + split passes.  Every group runs in its own subprocess under a
timeout, so a trapped kernel cannot take the other groups (or the box) down.

    python scripts/gemm_f16_check.py                  # all groups -> JSON lines
    python scripts/gemm_f16_check.py --one fwd 3      # one group in this process
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "surface-texture-inpainting-net_b200"))

SHAPES = [  # (M, N, K)
    (300, 128, 64), (1000, 192, 96), (129, 64, 32), (77, 36, 12),
    (327696, 256, 64), (327696, 64, 128), (81936, 512, 128), (81936, 128, 256), (20496, 1024, 256), (20496, 256, 512),
    (5136, 2048, 512), (5136, 512, 1024), (1296, 4096, 1024), (1296, 1024, 2048), (1296, 1024, 512),
]
# data regimes: (name, scale of x, scale of dy, heavy tails)
REGIMES = [("unit", 1.0, 1.0, False), ("tiny_grad", 3.0, 1e-9, False), ("tails", 1.0, 1e-4, True)]


def timeit(torch, fn, reps=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def one(op, passes, regime):
    import torch
    from stinet_b200 import _abi, ops
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    rname, sx, sdy, tails = [r for r in REGIMES if r[0] == regime][0]
    for (M, N, K) in SHAPES:
        g = torch.Generator(device="cpu").manual_seed(M + 7 * N + 13 * K)
        x = torch.randn(M, K, generator=g) * sx + 0.5 * sx
        w = torch.randn(N, K, generator=g) / K ** 0.5
        dy = torch.randn(M, N, generator=g) * sdy
        if tails:
            x = x * torch.exp(3 * torch.randn(M, K, generator=g))
            dy = dy * torch.exp(3 * torch.randn(M, N, generator=g))
        x, w, dy = x.to(dev), w.to(dev), dy.to(dev)
        b = torch.randn(N, generator=g).to(dev) * sx
        mask = (torch.rand(M, generator=g) > 0.1).to(torch.int32).to(dev)
        nb = _abi.query("stinet_gemm_workspace_bytes", M, N, K, 0)
        ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
        need_lo = passes == 3
        t_split = timeit(torch, lambda: (ops.invalidate_planes(), ops.planes_of(x, need_lo)))
        xp, wp, dyp = ops.planes_of(x, need_lo), ops.planes_of(w, need_lo), ops.planes_of(dy, need_lo)
        P = ops._ptr
        if op == "fwd":
            out = torch.full((M, N), float("nan"), device=dev)
            def run():
                _abi.call("stinet_linear_fwd_f16", xp.hi.data_ptr(), P(xp.lo), xp.ld, xp.exp.data_ptr(), wp.hi.data_ptr(),
                          P(wp.lo), wp.ld, wp.exp.data_ptr(), b.data_ptr(), mask.data_ptr(), out.data_ptr(), N, None, M, N, K,
                          passes, ws.data_ptr(), nb, stream)
            def old():
                _abi.call("stinet_linear_fwd", x.data_ptr(), K, w.data_ptr(), K, b.data_ptr(), mask.data_ptr(),
                          out.data_ptr(), N, M, N, K, 0, ws.data_ptr(), nb, stream)
            ref = x.double() @ w.double().t() + b.double() * (mask > 0).double().unsqueeze(1)
        elif op == "dgrad":
            out = torch.full((M, K), float("nan"), device=dev)
            def run():
                _abi.call("stinet_linear_dgrad_f16", dyp.hi.data_ptr(), P(dyp.lo), dyp.ld, dyp.exp.data_ptr(),
                          wp.hi.data_ptr(), P(wp.lo), wp.ld, wp.exp.data_ptr(), out.data_ptr(), K, None, M, N, K, passes,
                          ws.data_ptr(), nb, stream)
            def old():
                _abi.call("stinet_linear_dgrad", dy.data_ptr(), N, w.data_ptr(), K, out.data_ptr(), K, M, N, K, 0,
                          ws.data_ptr(), nb, stream)
            ref = dy.double() @ w.double()
        else:
            out = torch.full((N, K), float("nan"), device=dev)
            def run():
                _abi.call("stinet_linear_wgrad_f16", dyp.hi.data_ptr(), P(dyp.lo), dyp.ld, dyp.exp.data_ptr(),
                          xp.hi.data_ptr(), P(xp.lo), xp.ld, xp.exp.data_ptr(), out.data_ptr(), K, M, N, K, passes,
                          ws.data_ptr(), nb, stream)
            def old():
                _abi.call("stinet_linear_wgrad", dy.data_ptr(), N, x.data_ptr(), K, None, out.data_ptr(), K, None, M, N, K,
                          0, ws.data_ptr(), nb, stream)
            ref = dy.double().t() @ x.double()
        rec = {"op": op, "passes": passes, "regime": rname, "M": M, "N": N, "K": K}
        run()
        torch.cuda.synchronize()
        rec["err"] = float((out.double() - ref).abs().max() / ref.abs().max())
        first = out.clone()
        run()
        torch.cuda.synchronize()
        rec["deterministic"] = bool(torch.equal(first, out))
        rec["ms"] = timeit(torch, run)
        rec["TFLOPs"] = 2.0 * M * N * K / (rec["ms"] * 1e-3) / 1e12
        rec["ms_amax_split_x"] = t_split
        if (K % 4 == 0 and N % 4 == 0):
            old()
            torch.cuda.synchronize()
            rec["err_tf32x3"] = float((out.double() - ref).abs().max() / ref.abs().max())
            rec["ms_tf32x3"] = timeit(torch, old)
        # a loose amax bound (2^10 above the true maximum) must not cost accuracy
        if passes == 3 and op == "fwd":
            xb = x.clone()
            ops.set_amax(xb, x.abs().max().reshape(1) * 1024.0)
            xq = ops.planes_of(xb, True)
            _abi.call("stinet_linear_fwd_f16", xq.hi.data_ptr(), P(xq.lo), xq.ld, xq.exp.data_ptr(), wp.hi.data_ptr(),
                      P(wp.lo), wp.ld, wp.exp.data_ptr(), b.data_ptr(), mask.data_ptr(), out.data_ptr(), N, None, M, N, K,
                      passes, ws.data_ptr(), nb, stream)
            torch.cuda.synchronize()
            rec["err_loose_bound"] = float((out.double() - ref).abs().max() / ref.abs().max())
        print(json.dumps(rec), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", nargs=3, default=None)
    ap.add_argument("--regimes", default="unit,tiny_grad,tails")
    a = ap.parse_args()
    if a.one:
        one(a.one[0], int(a.one[1]), a.one[2])
        return
    for regime in a.regimes.split(","):
        for passes in (3, 1):
            if passes == 1 and regime != "unit":
                continue
            for op in ("fwd", "dgrad", "wgrad"):
                try:
                    r = subprocess.run([sys.executable, __file__, "--one", op, str(passes), regime], timeout=300,
                                       capture_output=True, text=True)
                    sys.stdout.write(r.stdout)
                    if r.returncode != 0:
                        print(json.dumps({"op": op, "passes": passes, "regime": regime, "FAILED": r.returncode,
                                          "stderr": r.stderr[-1500:]}), flush=True)
                except subprocess.TimeoutExpired:
                    print(json.dumps({"op": op, "passes": passes, "regime": regime, "FAILED": "timeout"}), flush=True)


if __name__ == "__main__":
    main()
