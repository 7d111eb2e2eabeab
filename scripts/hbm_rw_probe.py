"""Write-only, read-only and copy bandwidth of the HBM on this box (torch fill_/sum/copy_ on 2 GiB, CUDA events):
which roofline a kernel that mostly writes (GEMM epilogues with tiny K, unpool) can be held against."""
import json
import torch

def t(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

n = 1 << 29                                   # 2 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
a.normal_()
out = {"bytes": 4 * n}
out["write_GBps"] = 4 * n / t(lambda: b.fill_(1.5)) / 1e9
out["memset_GBps"] = 4 * n / t(lambda: b.zero_()) / 1e9
out["read_GBps"] = 4 * n / t(lambda: a.sum()) / 1e9
out["copy_GBps"] = 8 * n / t(lambda: b.copy_(a)) / 1e9
for mb in (64, 256):                            # the sizes of one level-0 activation: L2 write-back effects
    m = mb * (1 << 20) // 4
    out[f"write_{mb}MB_GBps"] = 4 * m / t(lambda: b[:m].fill_(1.5), 50) / 1e9
print(json.dumps(out))
