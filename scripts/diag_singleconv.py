"""SingleConvMeshNet on the B200, every tensor reported (no early stop): forward / loss, every gradient against the golden
vectors AND against the fp64 oracle replaying the CUDA path's decisions (next to the fp32 oracle's own distance from that
truth), BatchNorm buffers after one training step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "surface-texture-inpainting-net_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from conftest import rel_err
from test_singleconv import FIXTURES, load, record_product_decisions, oracle_with_replayed_decisions
from stinet_b200.models.singleconvmeshnet import SingleConvMeshNet
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
for name in FIXTURES:
    fix = load(name)
    net = SingleConvMeshNet(**fix["kwargs"], precision=prec)
    net.load_state_dict(fix["state_dict"], strict=True)
    net = net.to("cuda").train()
    b = fix["batch"].to("cuda")
    choices = record_product_decisions(net, b)
    b.x = b.x.detach().clone().requires_grad_(True)
    out = net(b)
    loss = out.square().mean()
    print(f"== {name} [{prec}] out {rel_err(out, fix['out']):.2e} loss {rel_err(loss, fix['loss']):.2e}", flush=True)
    if fix.get("forward_only"):
        continue
    loss.backward()
    t_out, t_loss, t_grads, dec = oracle_with_replayed_decisions(fix, choices)
    r_out, r_loss, r_grads, _ = oracle_with_replayed_decisions(fix, choices, torch.float32)
    print(f"   vs replay truth: out {rel_err(out, t_out):.2e} (fp32 oracle {rel_err(r_out, t_out):.2e}); decisions differing: relu {dec.n_relu_diff} "
          f"(margin {dec.max_relu_margin:.1e}) pool {dec.n_pool_diff} (margin {dec.max_pool_margin:.1e})")
    got = {k: p.grad for k, p in net.named_parameters()}
    got["__x__"] = b.x.grad
    gold = dict(fix["grads"], __x__=fix["grad_x"])
    scale = max(float(v.abs().max()) for v in t_grads.values())
    for k in t_grads:
        nb = float(t_grads[k].abs().max())
        tag = "tiny" if nb < 1e-4 * scale else ""
        print(f"   {k:55s} |g| {nb:.1e} {tag:4s} cuda-vs-truth {rel_err(got[k], t_grads[k]):.2e}  fp32oracle-vs-truth {rel_err(r_grads[k], t_grads[k]):.2e}  "
              f"cuda-vs-golden {rel_err(got[k], gold[k]):.2e}  golden-vs-truth {rel_err(gold[k], t_grads[k]):.2e}")
    for k, v in net.named_buffers():
        ref = fix["buffers_after"][k]
        if ref.is_floating_point():
            print(f"   buffer {k:48s} abs err {float((v.detach().cpu() - ref).abs().max()):.2e} (|ref| {float(ref.abs().max()):.1e})")
