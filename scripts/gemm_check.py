"""GPU check of the dense-layer entry points (stinet_linear_{fwd,dgrad,wgrad}) in every precision mode against an
fp64 torch reference: max-norm relative error and CUDA-event timing per shape.  Each (op, precision) group runs in
its own subprocess under a timeout, so a trapped kernel cannot take the other groups (or the box) down.

    python scripts/gemm_check.py                 # all groups
    python scripts/gemm_check.py --one fwd fp32  # one group in this process
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "surface-texture-inpainting-net_b200"))

SHAPES = [  # (M, N, K)
    (300, 128, 64), (1000, 192, 96), (129, 64, 32), (5136, 512, 1024),
    (327696, 256, 64), (327696, 64, 128), (81936, 512, 128), (1296, 4096, 1024), (1296, 1024, 2048),
    (5136, 2048, 512), (5136, 512, 1024), (20496, 1024, 256), (20496, 256, 512), (1296, 1024, 512), (81936, 128, 256),
]


SMALL = [(8192, 128, 4), (8192, 32, 64), (8192, 32, 4), (2048, 256, 128), (2048, 64, 128), (2048, 64, 32),
         (512, 512, 128), (512, 128, 256), (7686, 32, 16), (7686, 16, 32), (1926, 128, 32), (486, 256, 64),
         (8192, 64, 64), (8192, 32, 32), (126, 512, 128), (126, 128, 256)]


def one(op, prec):
    import torch
    from stinet_b200 import _abi
    from stinet_b200._abi import PREC
    dev = torch.device("cuda", 0)
    p = PREC[prec]
    stream = torch.cuda.current_stream().cuda_stream
    out = []
    for (M, N, K) in (SMALL if os.environ.get("GEMM_CHECK_SMALL") else SHAPES):
        g = torch.Generator(device="cpu").manual_seed(M + 7 * N + 13 * K)
        x = torch.randn(M, K, generator=g).to(dev)
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
        b = torch.randn(N, generator=g).to(dev)
        dy = torch.randn(M, N, generator=g).to(dev)
        mask = (torch.rand(M, generator=g) > 0.1).to(torch.int32).to(dev)
        nb = _abi.query("stinet_gemm_workspace_bytes", M, N, K, p)
        ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
        if op == "fwd":
            y = torch.full((M, N), float("nan"), device=dev)
            def run():
                _abi.call("stinet_linear_fwd", x.data_ptr(), K, w.data_ptr(), K, b.data_ptr(), mask.data_ptr(),
                          y.data_ptr(), N, M, N, K, p, ws.data_ptr(), nb, stream)
            ref = x.double() @ w.double().t() + b.double() * (mask > 0).double().unsqueeze(1)
            got = lambda: y
        elif op == "dgrad":
            dx = torch.full((M, K), float("nan"), device=dev)
            def run():
                _abi.call("stinet_linear_dgrad", dy.data_ptr(), N, w.data_ptr(), K, dx.data_ptr(), K, M, N, K, p,
                          ws.data_ptr(), nb, stream)
            ref = dy.double() @ w.double()
            got = lambda: dx
        else:
            dw = torch.full((N, K), float("nan"), device=dev)
            db = torch.full((N,), float("nan"), device=dev)
            def run():
                _abi.call("stinet_linear_wgrad", dy.data_ptr(), N, x.data_ptr(), K, mask.data_ptr(), dw.data_ptr(), K,
                          db.data_ptr(), M, N, K, p, ws.data_ptr(), nb, stream)
            ref = dy.double().t() @ x.double()
            got = lambda: dw
        run()
        torch.cuda.synchronize()
        err = float((got().double() - ref).abs().max() / ref.abs().max())
        extra = {}
        if op == "wgrad":
            rb = (dy.double() * (mask > 0).double().unsqueeze(1)).sum(0)
            extra["err_dbias"] = float((db.double() - rb).abs().max() / rb.abs().max())
        first = got().clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            run()
        reps = 10
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rec = {"op": op, "prec": prec, "M": M, "N": N, "K": K, "rel_err": err, "ms": round(ms, 4),
               "TFLOPs": round(2 * M * N * K / ms / 1e9, 2), "deterministic": bool(torch.equal(first, got())), **extra}
        print(json.dumps(rec), flush=True)
        out.append(rec)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", nargs=2)
    ap.add_argument("--precs", default="fp32,tf32,bf16,fp32_simt")
    ap.add_argument("--ops", default="fwd,dgrad,wgrad")
    a = ap.parse_args()
    if a.one:
        one(*a.one)
        return
    for prec in a.precs.split(","):
        for op in a.ops.split(","):
            try:
                r = subprocess.run([sys.executable, __file__, "--one", op, prec], timeout=300, capture_output=True, text=True)
                sys.stdout.write(r.stdout)
                if r.returncode != 0:
                    print(json.dumps({"op": op, "prec": prec, "FAILED": r.returncode, "stderr": r.stderr[-800:]}), flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps({"op": op, "prec": prec, "FAILED": "timeout"}), flush=True)


if __name__ == "__main__":
    main()
