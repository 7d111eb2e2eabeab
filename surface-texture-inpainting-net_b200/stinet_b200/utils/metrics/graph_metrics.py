"""Mirror of reference utils/metrics/graph_metrics.py:6-72 -- the metrics the 3D trainer evaluates after EVERY
training / validation step (trainers/inpainting3d_trainer.py:254-263) -- on the sm_100a kernels of metrics.cu.

Same names and call signatures (`GraphLaplaceOperator()(x, edge_index)`, `GraphLaplaceVariance()(x, edge_indices)`,
`graph_total_variation(x, edge_indices)`, `psnr(x, y, data_range)`).  `edge_index` may be the reference's [2, E]
int64 tensor (a CSR is built on the fly) or the `EdgeCSR` the forward pass already holds
(`GraphCache.for_sample(sample, L).edges('edge_index', 0)`), which is how the per-step call avoids any sort.
`psnr` takes an optional `mask`: `psnr(x, y, 2.0, mask=m)` equals the reference's `psnr(x[m > 0], y[m > 0], 2.0)`
without the boolean-index copies (and without the host sync their data-dependent shape costs).
Results are 0-dim / [1] device tensors like the reference's; there is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Union

import torch

from ... import _abi
from ...graph import EdgeCSR, _ptr, _stream
from ...models.modules._structure import as_edge_csr
from ...ops import _ld, _mat


def _ws(n: int, device):
    nb = _abi.query("stinet_metrics_workspace_bytes", n)
    return torch.empty(max(nb, 16), dtype=torch.uint8, device=device), nb


class GraphLaplaceOperator(torch.nn.Module):
    """out_i = sum_{j->i} x_j - deg_i * x_i   (reference :6-16, MessagePassing(aggr='add'))"""

    def forward(self, x, edge_index):
        x = _mat(x)
        n, c = x.shape
        csr = as_edge_csr(edge_index, n)
        out = torch.empty((n, c), dtype=torch.float32, device=x.device)
        _abi.call("stinet_graph_laplace", x.data_ptr(), _ld(x), csr.rowptr_t.data_ptr(), csr.col_t.data_ptr(), n, c,
                  out.data_ptr(), c, _stream(), cost=(csr.e * (4 * c + 4) + n * (8 * c + 4), 0, f"C{c}"))
        return out


class GraphLaplaceVariance(torch.nn.Module):
    """biased variance over the vertices of the Laplacian of the grey channel (reference :19-30) -> tensor [1]"""

    def __init__(self):
        super().__init__()
        self.filter = GraphLaplaceOperator()

    def grayscale(self, x):
        return 0.299 * x[:, 0:1] + 0.587 * x[:, 1:2] + 0.114 * x[:, 2:3]

    def forward(self, x, edge_indices):
        x = _mat(x)
        n, c = x.shape
        if c < 3:
            raise _abi.StinetError("GraphLaplaceVariance expects RGB rows")
        csr = as_edge_csr(edge_indices, n)
        out = torch.empty((1,), dtype=torch.float32, device=x.device)
        ws, nb = _ws(n, x.device)
        _abi.call("stinet_graph_laplace_variance", x.data_ptr(), _ld(x), csr.rowptr_t.data_ptr(),
                  csr.col_t.data_ptr(), n, out.data_ptr(), ws.data_ptr(), nb, _stream(),
                  cost=(csr.e * (12 + 4) + n * (12 + 4), 0, ""))
        return out


def graph_total_variation(x, edge_indices):
    """sum_e |x[src] - x[dst]| / (N * C)   (reference :33-37) -> 0-dim tensor"""
    x = _mat(x)
    n, c = x.shape
    csr = as_edge_csr(edge_indices, n)
    out = torch.empty((1,), dtype=torch.float32, device=x.device)
    ws, nb = _ws(n, x.device)
    _abi.call("stinet_graph_total_variation", x.data_ptr(), _ld(x), csr.rowptr_t.data_ptr(), csr.col_t.data_ptr(), n, c,
              out.data_ptr(), ws.data_ptr(), nb, _stream(), cost=(csr.e * (4 * c + 4) + n * (4 * c + 4), 0, f"C{c}"))
    return out[0]


def psnr(x: torch.Tensor, y: torch.Tensor, data_range: Union[int, float] = 1.0, convert_to_greyscale: bool = False,
         mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """-10 log10(mean(((x - y) / data_range)^2) + 1e-8)   (reference :40-72) -> 0-dim tensor.
    `mask` [N] or [N,1]: restrict to rows with mask > 0 (the trainer's `psnr_mask_only`)."""
    if convert_to_greyscale:
        raise NotImplementedError("convert_to_greyscale is image-only code in the reference (a 4-D view of a [N,3] "
                                  "tensor, :64) and is never used by the graph trainers")
    x, y = _mat(x), _mat(y)
    n, c = x.shape
    assert y.shape == x.shape
    m = None
    if mask is not None:
        m = mask.reshape(-1).to(torch.float32).contiguous()
        assert m.numel() == n
    out = torch.empty((2,), dtype=torch.float32, device=x.device)
    ws, nb = _ws(n, x.device)
    _abi.call("stinet_psnr", x.data_ptr(), _ld(x), y.data_ptr(), _ld(y), _ptr(m), n, c, float(data_range), out.data_ptr(),
              ws.data_ptr(), nb, _stream(), cost=(8 * n * c + (4 * n if m is not None else 0), 0, f"C{c}"))
    return out[0]
