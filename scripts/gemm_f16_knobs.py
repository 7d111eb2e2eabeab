"""One process = one setting of the planes-GEMM knobs (STINET_TC_PROMOTE16, STINET_TC_CORR_ONCE are read once per process):
error vs fp64, CUDA-event time and -- with the debug library (STINET_B200_LIB=...libstinet_b200_dbg.so) -- the per-role
barrier-wait counters of CTA 0 for fwd / dgrad / wgrad on the shapes of BASELINE config 2."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "surface-texture-inpainting-net_b200"))
import torch
from stinet_b200 import _abi, ops
lib = _abi.load()
dbg = hasattr(lib, "stinet_tc_debug_read")
if dbg:
    lib.stinet_tc_debug_read.argtypes = [ctypes.c_void_p]
names = ["prod<-empty", "split<-full", "mma<-operands", "mma<-tmem_empty", "epi<-tmem_full", "epi_store", "total", "split_work", "units"]
dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
P = ops._ptr
tag = {"promote16": os.environ.get("STINET_TC_PROMOTE16", "2"), "corr_once": os.environ.get("STINET_TC_CORR_ONCE", "0"), "dbg": dbg}
SHAPES = [(327696, 256, 64), (327696, 64, 128), (81936, 512, 128), (20496, 1024, 256), (20496, 256, 512), (5136, 2048, 512),
          (5136, 512, 1024), (1296, 4096, 1024), (1296, 1024, 2048)]
for (M, N, K) in SHAPES:
    g = torch.Generator(device="cpu").manual_seed(M + 7 * N + 13 * K)
    x = (torch.randn(M, K, generator=g) + 0.5).to(dev); w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    dy = torch.randn(M, N, generator=g).to(dev)
    xp, wp, dyp = ops.planes_of(x), ops.planes_of(w), ops.planes_of(dy)
    nb = _abi.query("stinet_gemm_workspace_bytes", M, N, K, 0); ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
    for op in ("fwd", "dgrad", "wgrad"):
        if op == "fwd":
            out = torch.empty(M, N, device=dev)
            run = lambda: _abi.call("stinet_linear_fwd_f16", xp.hi.data_ptr(), P(xp.lo), xp.ld, xp.exp.data_ptr(), wp.hi.data_ptr(), P(wp.lo), wp.ld,
                                    wp.exp.data_ptr(), None, None, out.data_ptr(), N, None, M, N, K, 3, ws.data_ptr(), nb, st)
            ref = lambda: x.double() @ w.double().t()
        elif op == "dgrad":
            out = torch.empty(M, K, device=dev)
            run = lambda: _abi.call("stinet_linear_dgrad_f16", dyp.hi.data_ptr(), P(dyp.lo), dyp.ld, dyp.exp.data_ptr(), wp.hi.data_ptr(), P(wp.lo), wp.ld,
                                    wp.exp.data_ptr(), out.data_ptr(), K, None, M, N, K, 3, ws.data_ptr(), nb, st)
            ref = lambda: dy.double() @ w.double()
        else:
            out = torch.empty(N, K, device=dev)
            run = lambda: _abi.call("stinet_linear_wgrad_f16", dyp.hi.data_ptr(), P(dyp.lo), dyp.ld, dyp.exp.data_ptr(), xp.hi.data_ptr(), P(xp.lo), xp.ld,
                                    xp.exp.data_ptr(), out.data_ptr(), K, M, N, K, 3, ws.data_ptr(), nb, st)
            ref = lambda: dy.double().t() @ x.double()
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        r = ref()
        rec = dict(tag, op=op, M=M, N=N, K=K, err=float((out.double() - r).abs().max() / r.abs().max()))
        del r
        if dbg:
            buf = (ctypes.c_ulonglong * 16)()
            assert lib.stinet_tc_debug_read(buf) == 0
            rec["waits"] = {n: int(buf[i]) for i, n in enumerate(names)}
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run()
            e1.record(); torch.cuda.synchronize()
            rec["ms"] = e0.elapsed_time(e1) / 10
            rec["TFLOPs"] = 2.0 * M * N * K / (rec["ms"] * 1e-3) / 1e12
        print(json.dumps(rec), flush=True)
