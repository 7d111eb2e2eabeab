"""Per-tensor error of the CUDA path against one golden fixture, for the tensor-core and FFMA dense-layer modes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "surface-texture-inpainting-net_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import load_golden, rel_err
from test_gpu_model import _net_from, _loss
name = sys.argv[1] if len(sys.argv) > 1 else "edgeconvtransinv_ico_b2"
fix = load_golden(name)
for prec in ("fp32", "fp32_simt"):
    net = _net_from(fix["kwargs"], fix["state_dict"]); net.set_precision(prec)
    batch = fix["batch"].to("cuda"); batch.x = batch.x.clone().requires_grad_(True)
    out = net(batch); loss = _loss(out, batch); loss.backward()
    print(prec, "out", "%.2e" % rel_err(out, fix["out"]), "loss %.2e" % rel_err(loss, fix["loss"]), "gx %.2e" % rel_err(batch.x.grad, fix["grad_x"]))
    scale = max(float(v.abs().max()) for v in fix["grads"].values())
    for k, p in net.named_parameters():
        b = fix["grads"][k]
        if float(b.abs().max()) < 1e-4 * scale: continue
        e = rel_err(p.grad, b)
        if e > 3e-6: print("   ", k, tuple(b.shape), "%.2e" % e)
