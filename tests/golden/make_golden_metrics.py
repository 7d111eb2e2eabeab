"""Golden vectors for the per-step graph metrics, minted from the reference's OWN
utils/metrics/graph_metrics.py (imported unmodified; PyG's MessagePassing comes from tests/golden/pyg_shim).

    python tests/golden/make_golden_metrics.py     # rewrites tests/golden/metrics/graph_metrics.pt

Inputs follow the call site trainers/inpainting3d_trainer.py:254-263: prediction / ground truth [N,3] in (-1,1),
mask [N,1], the level-0 edge_index; psnr with data_range=2.0, once over all vertices and once over the masked ones.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "surface-texture-inpainting-net_b200"))
sys.path.insert(0, os.path.join(HERE, "pyg_shim"))
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402
from utils.metrics import graph_metrics as ref  # noqa: E402  (REAL reference code)
from stinet_b200 import synthetic  # noqa: E402


def main():
    cases = {}
    specs = {"ico3": ("icosphere", dict(subdiv=3, n_levels=1, seed=61, mask_radius=3)),
             "grid12": ("grid", dict(size=12, n_levels=1, seed=62)),
             "graph18": None}
    for name, spec in specs.items():
        g = torch.Generator().manual_seed(len(name))
        if spec is None:
            ei, n = synthetic.paper_graph18()            # irregular degrees, one isolated vertex
            mask = (torch.rand(n, 1, generator=g) > 0.5).float()
        else:
            s = {"icosphere": synthetic.icosphere_sample, "grid": synthetic.grid_sample}[spec[0]](**spec[1])
            ei, n, mask = s.edge_index, s.num_nodes, s.mask.float()
        pred = torch.rand(n, 3, generator=g) * 2 - 1
        gt = torch.rand(n, 3, generator=g) * 2 - 1
        lapvar = ref.GraphLaplaceVariance()
        sel = mask.squeeze() > 0
        cases[name] = {
            "edge_index": ei.clone(), "n": n, "pred": pred, "gt": gt, "mask": mask,
            "laplace": ref.GraphLaplaceOperator()(pred, ei),
            "lap_var": lapvar(pred, ei),
            "tv": ref.graph_total_variation(pred, ei),
            "psnr": ref.psnr(pred, gt, data_range=2.0),
            "psnr_mask_only": ref.psnr(pred[sel], gt[sel], data_range=2.0),
            "psnr_grey_range1": ref.psnr(pred, gt, data_range=1.0, convert_to_greyscale=False),
        }
        print(name, n, {k: (v.tolist() if torch.is_tensor(v) and v.numel() < 4 else None) for k, v in cases[name].items()
                        if k in ("lap_var", "tv", "psnr", "psnr_mask_only")})
    torch.save(cases, os.path.join(HERE, "metrics", "graph_metrics.pt"))


if __name__ == "__main__":
    main()
