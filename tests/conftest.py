import glob
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "surface-texture-inpainting-net_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    from stinet_b200.data import GraphBatch
    fix = torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)
    fix["batch"] = GraphBatch(**fix["sample"])
    return fix


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """SURVEY 8c tolerance definition: ||a-b||_inf / max(||b||_inf, tiny), per tensor."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def assert_grads_close(got: dict, ref: dict, tol: float, what: str = "", slack: dict = None):
    """Per-tensor ||a-b||inf / ||b||inf <= tol for every gradient tensor.  Gradients that are structurally zero in
    exact arithmetic (e.g. the bias feeding an instance norm: the norm removes any constant) hold nothing but rounding
    noise on both sides, so a relative comparison is meaningless there: a tensor whose reference norm is below 1e-4 of
    the largest gradient norm in the set only has to be equally negligible."""
    scale = max(float(v.detach().abs().max()) for v in ref.values())
    for k, b in ref.items():
        a = got[k]
        nb = float(b.detach().abs().max())
        if nb < 1e-4 * scale:
            assert float(a.detach().abs().max()) < 1e-4 * scale, f"{what}{k}: expected a negligible gradient"
        else:
            e = rel_err(a, b)
            t = tol + (min(slack.get(k, 0.0), tol) if slack else 0.0)     # slack: the reference side's own error, capped
            assert e <= t, f"{what}{k}: rel err {e:.3e} > {t:.3e}"


class cuda_decisions:
    """Context manager that records, in call order, the discrete choices the CUDA path takes during one forward pass:
    ("relu", bool [E, 2*dout]) per fused EdgeConv message stage -- the sign of P[target] + Q[source] evaluated with the
    very fp32 addition the kernel performs, in ORIGINAL edge order -- and ("pool", long [n_coarse, C]) per max-pool.
    Feed `.choices` to oracle.Decisions.replay (see its docstring for why)."""

    def __enter__(self):
        from stinet_b200 import ops
        self.ops, self.choices = ops, []
        self._em, self._pm = ops.edge_message, ops.pool_max

        def edge_message(pq, csr):
            h = pq.shape[1] // 2
            with torch.no_grad():
                m = (pq[csr._dst, :h] + pq[csr._src, h:]) > 0
            self.choices.append(("relu", m.cpu()))
            return self._em(pq, csr)

        def pool_max(x, cl):
            out, arg = self._pm(x, cl)
            self.choices.append(("pool", arg.detach().long().cpu()))
            return out, arg

        def observer(pq, csr):                       # the fused EdgeConv node (ops.EdgeConvFn) reports its [P | Q] here
            h = pq.shape[1] // 2
            with torch.no_grad():
                m = (pq[csr._dst, :h] + pq[csr._src, h:]) > 0
            self.choices.append(("relu", m.cpu()))

        ops.edge_message, ops.pool_max = edge_message, pool_max
        ops._edge_message_observer = observer
        return self

    def __exit__(self, *exc):
        self.ops.edge_message, self.ops.pool_max = self._em, self._pm
        self.ops._edge_message_observer = None
