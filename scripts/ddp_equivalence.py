"""Run under torchrun with N >= 2 ranks (one per GPU, NCCL):
  (1) the reducer's averaged gradients of the real STINet == the single-process gradient of the mean loss over all ranks'
      batches (every rank recomputes that reference itself on its own GPU), to fp32 tolerance;
  (2) GraphedTrainStep with the all-reduce captured inside the graph == the eager multi-rank step, bit for bit, over a few
      optimizer steps (losses and final parameters);
  (3) every rank ends with identical parameters.
Prints DDP_EQUIVALENCE_OK from rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "surface-texture-inpainting-net_b200"), ROOT):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
from stinet_b200 import synthetic
from stinet_b200.engine import GraphedTrainStep
from stinet_b200.models import surfacetextureinpaintingnet as S
from stinet_b200.parallel import GradAllReducer, init_distributed


def loss_fn(out, b):
    composed = torch.where((b.mask > 0).expand_as(b.color), out, b.color)
    return ((composed - b.color).abs() * torch.pow(0.99, b.mask.squeeze().float()).unsqueeze(1)).mean()


def make_net(dev):
    torch.manual_seed(49)
    return S.define_G(input_nc=10, output_nc=3, ngf=32, filter_type="edgeconvtransinv", norm="instance", n_blocks=3, n_levels=2,
                      pooling_type="max", gpu_ids=[dev]).train()


def main():
    rank, local, world = init_distributed()
    assert world >= 2
    dev = torch.device("cuda", local)
    batches = [[synthetic.make_batch("icosphere", 2, 2, seed=100 + 10 * r + s, subdiv=4, mask_radius=4) for s in range(3)]
               for r in range(world)]
    # ---- (1) averaged gradients == single-process gradient of the mean loss
    net = make_net(dev)
    red = GradAllReducer(net, bucket_bytes=256 << 10)
    assert len(red.buckets) > 2
    b = batches[rank][0].to(dev)
    red.zero_grad()
    (loss_fn(net(b), b) * red.loss_scale).backward()
    red.finish()
    got = [p.grad.detach().clone() for p in net.parameters()]
    ref_net = make_net(dev)
    total = 0
    for r in range(world):
        rb = batches[r][0].to(dev)
        total = total + loss_fn(ref_net(rb), rb) / world
    total.backward()
    scale = max(float(p.grad.abs().max()) for p in ref_net.parameters())
    for g, p in zip(got, ref_net.parameters()):
        if float(p.grad.abs().max()) < 1e-4 * scale:
            continue
        err = float((g - p.grad).abs().max() / p.grad.abs().max())
        assert err <= 1e-5, err
    # ---- (2) graphed step (collectives inside the graph) == eager step
    def run(graphed):
        net = make_net(dev)
        red = GradAllReducer(net, bucket_bytes=256 << 10)
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, amsgrad=True, fused=True, capturable=True)
        losses = []
        if graphed:
            step = GraphedTrainStep(net, loss_fn, opt, red, warmup=1)
            for hb in batches[rank]:
                losses.append(float(step(hb.pin_memory()).item()))
        else:
            for hb in batches[rank]:
                gb = hb.to(dev)
                red.zero_grad()
                loss = loss_fn(net(gb), gb)
                (loss * red.loss_scale).backward()
                red.finish()
                opt.step()
                losses.append(float(loss.item()))
        return losses, [p.detach().clone() for p in net.parameters()]
    l_e, p_e = run(False)
    l_g, p_g = run(True)
    assert l_e == l_g, (l_e, l_g)
    for a, c in zip(p_e, p_g):
        assert torch.equal(a, c)
    # ---- (3) replicas stay identical
    for p in p_g:
        ref = p.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, p)
    dist.barrier()
    if rank == 0:
        print("DDP_EQUIVALENCE_OK", world, flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)         # graphs with captured NCCL kernels are alive: skip the (possibly blocking) communicator teardown


if __name__ == "__main__":
    main()
