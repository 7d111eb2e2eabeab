"""torch_scatter 2.0.x, CPU semantics, restated in plain torch (test-only).

scatter(src, index, dim=0, dim_size, reduce): 'sum'/'add', 'mean' (= sum / clamp(count, 1)), 'max', 'min'.
scatter_max returns (out, arg): strict '>' update in index order => FIRST occurrence wins ties;
segments that receive nothing have out = 0 and arg = src.size(dim); backward routes grad to arg only.
"""
import torch


def _expand_index(index, src):
    if index.dim() == src.dim():
        return index
    shape = [-1] + [1] * (src.dim() - 1)
    return index.view(*shape).expand_as(src)


def scatter_sum(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0 and out is None
    n = int(dim_size) if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add_(0, index, src)


scatter_add = scatter_sum


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    s = scatter_sum(src, index, dim, out, dim_size)
    cnt = scatter_sum(torch.ones(index.size(0), dtype=src.dtype, device=src.device), index, 0, None, s.size(0))
    cnt = cnt.clamp_(min=1).view([-1] + [1] * (src.dim() - 1))
    if src.is_floating_point():
        return s / cnt
    return torch.div(s, cnt, rounding_mode="floor")


class _ScatterMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, n):
        n_src = src.size(0)
        idx = _expand_index(index, src)
        lowest = torch.finfo(src.dtype).min if src.is_floating_point() else torch.iinfo(src.dtype).min
        out = torch.full((n,) + tuple(src.shape[1:]), lowest, dtype=src.dtype, device=src.device)
        out = out.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
        # first row (lowest position in index order) attaining the segment maximum
        pos = torch.arange(n_src, device=src.device).view([-1] + [1] * (src.dim() - 1)).expand_as(src)
        cand = torch.where(src == out.gather(0, idx), pos, torch.full_like(pos, n_src))
        arg = torch.full(out.shape, n_src, dtype=torch.long, device=src.device)
        arg = arg.scatter_reduce(0, idx, cand, reduce="amin", include_self=True)
        out = out.masked_fill(arg == n_src, 0)
        ctx.save_for_backward(arg)
        ctx.n_src = n_src
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, grad_out, _grad_arg):
        (arg,) = ctx.saved_tensors
        g = torch.zeros((ctx.n_src + 1,) + tuple(grad_out.shape[1:]), dtype=grad_out.dtype, device=grad_out.device)
        g.scatter_(0, arg, grad_out)
        return g[: ctx.n_src], None, None


def scatter_max(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0 and out is None
    n = int(dim_size) if dim_size is not None else int(index.max()) + 1
    return _ScatterMax.apply(src, index, n)


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == "max":
        return scatter_max(src, index, dim, out, dim_size)[0]
    raise ValueError(reduce)
