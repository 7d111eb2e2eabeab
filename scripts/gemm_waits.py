"""Where the persistent GEMM's warp roles wait (debug build with -DSTINET_TC_DEBUG, see gemm_tc.cu): cycles of CTA 0."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "surface-texture-inpainting-net_b200"))
import torch
from stinet_b200 import _abi
from stinet_b200._abi import PREC
lib = _abi.load()
lib.stinet_tc_debug_read.argtypes = [ctypes.c_void_p]
names = ["producer<-empty", "split<-full", "mma<-operands", "mma<-tmem_empty", "epi<-tmem_full", "epi store", "total", "split work", "units"]
dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
for prec in sys.argv[1:] or ["fp32", "tf32"]:
    p = PREC[prec]
    for (M, N, K) in [(327696, 256, 64), (327696, 64, 128), (81936, 512, 128), (1296, 4096, 1024), (1296, 1024, 2048), (5136, 512, 1024)]:
        x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev); y = torch.empty(M, N, device=dev)
        nb = _abi.query("stinet_gemm_workspace_bytes", M, N, K, p); ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
        for _ in range(3):
            _abi.call("stinet_linear_fwd", x.data_ptr(), K, w.data_ptr(), K, None, None, y.data_ptr(), N, M, N, K, p, ws.data_ptr(), nb, st)
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 16)()
        assert lib.stinet_tc_debug_read(buf) == 0
        tot = buf[6]
        print(prec, (M, N, K), " ".join(f"{n}={buf[i]}({100*buf[i]/max(tot,1):.0f}%)" if i not in (6, 8) else f"{n}={buf[i]}" for i, n in enumerate(names)))
