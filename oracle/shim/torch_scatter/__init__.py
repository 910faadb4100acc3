"""Test-side stand-in for `torch_scatter` (absent from this image, un-vendored and unpinned by the
reference: /root/reference/README.md:22, requirements.txt:1-4).

TEST INFRASTRUCTURE ONLY. It exists so that the UNMODIFIED reference (`/root/reference/mpqe/model.py:199`)
can be imported in this container to pin `oracle/mpqe_oracle.py` and to generate `tests/golden/`.
Semantics restated from the torch_scatter 2.x documentation for the call sites the reference uses
(`model.py:351-355, 381, 384, 509, 547`): reduce `src` rows into `out[index]` along `dim=0`,
`dim_size = index.max()+1` unless given.  `scatter_max/min` return `(values, arg)`; the tie-break contract
of this project is "smallest source row among the extrema" and empty segments give value 0 / arg = src.size(0).
"""
import torch


def _dim_size(index, dim_size):
    if dim_size is not None:
        return int(dim_size)
    return int(index.max().item()) + 1 if index.numel() else 0


def scatter_add(src, index, dim=0, out=None, dim_size=None, fill_value=0):
    assert dim == 0 and index.dim() == 1
    n = _dim_size(index, dim_size)
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add(0, index, src)


def scatter_mean(src, index, dim=0, out=None, dim_size=None, fill_value=0):
    n = _dim_size(index, dim_size)
    tot = scatter_add(src, index, dim, None, n)
    cnt = torch.zeros(n, dtype=src.dtype, device=src.device).index_add(
        0, index, torch.ones_like(index, dtype=src.dtype)).clamp(min=1)
    return tot / cnt.view((-1,) + (1,) * (src.dim() - 1))


def _scatter_extreme(src, index, dim_size, largest):
    assert index.dim() == 1
    n = _dim_size(index, dim_size)
    flat = src.reshape(src.size(0), -1)
    idx = index.view(-1, 1).expand_as(flat)
    init = torch.zeros((n, flat.size(1)), dtype=src.dtype, device=src.device)
    val = init.scatter_reduce(0, idx, flat, 'amax' if largest else 'amin', include_self=False)
    rows = torch.arange(src.size(0), device=src.device).view(-1, 1).expand_as(flat)
    hit = flat == val.index_select(0, index)
    cand = torch.where(hit, rows, torch.full_like(rows, src.size(0)))
    arg = torch.full((n, flat.size(1)), src.size(0), dtype=torch.long, device=src.device)
    arg = arg.scatter_reduce(0, idx, cand, 'amin', include_self=True)
    shape = (n,) + tuple(src.shape[1:])
    return val.reshape(shape), arg.reshape(shape)


def scatter_max(src, index, dim=0, out=None, dim_size=None, fill_value=None):
    assert dim == 0
    return _scatter_extreme(src, index, dim_size, True)


def scatter_min(src, index, dim=0, out=None, dim_size=None, fill_value=None):
    assert dim == 0
    return _scatter_extreme(src, index, dim_size, False)
