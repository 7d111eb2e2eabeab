"""A handful of dense-layer launches at the benchmark's shapes, for `ncu --set full -k regex:gemm_tc` captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "surface-texture-inpainting-net_b200"))
import torch
from stinet_b200 import _abi
from stinet_b200._abi import PREC
prec = PREC[sys.argv[1] if len(sys.argv) > 1 else "fp32"]
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
for (M, N, K) in [(327696, 256, 64), (1296, 4096, 1024), (81936, 512, 128)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    dy = torch.randn(M, N, device=dev); y = torch.empty(M, N, device=dev); dx = torch.empty(M, K, device=dev)
    dw = torch.empty(N, K, device=dev)
    nb = _abi.query("stinet_gemm_workspace_bytes", M, N, K, prec); ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
    for _ in range(2):
        _abi.call("stinet_linear_fwd", x.data_ptr(), K, w.data_ptr(), K, b.data_ptr(), None, y.data_ptr(), N, M, N, K, prec, ws.data_ptr(), nb, st)
        _abi.call("stinet_linear_dgrad", dy.data_ptr(), N, w.data_ptr(), K, dx.data_ptr(), K, M, N, K, prec, ws.data_ptr(), nb, st)
        _abi.call("stinet_linear_wgrad", dy.data_ptr(), N, x.data_ptr(), K, None, dw.data_ptr(), K, None, M, N, K, prec, ws.data_ptr(), nb, st)
    torch.cuda.synchronize()
