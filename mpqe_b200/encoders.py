"""Entity encoder: embedding lookup + L2 normalisation, fused in one CUDA kernel.

Mirror of the reference's `DirectEncoder` (mpqe/encoders.py:11-45): `enc(nodes, mode) -> [d, len(nodes)]` with
unit-norm columns (division by the norm, no eps).  The reference indexes `node_maps` on the host and copies the
rows' indices to the device on every call (mpqe/utils.py:22); here the id->row map lives on the device and the
gather, the norm and the division happen in `mpqe_gather_normalize_fwd`.
"""
import torch
import torch.nn as nn

from . import ops


class _GatherNormalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, id2row, ids, sparse_grad):
        with ops.device_guard(table.device):
            out = ops.gather_normalize(table, id2row, ids)
        ctx.save_for_backward(table, id2row, ids)
        ctx.sparse_grad = sparse_grad
        return out

    @staticmethod
    def backward(ctx, grad):
        table, id2row, ids = ctx.saved_tensors
        count = ids.numel()
        with ops.device_guard(table.device):
            rows = torch.empty(count, ops.D, dtype=torch.float32, device=table.device)
            rows_id = torch.empty(count, dtype=torch.int64, device=table.device)
            ops.gather_normalize_bwd(table, id2row, ids, grad.contiguous(), rows, rows_id)
            g = table_gradient(table, rows_id, rows, ctx.sparse_grad)
        return g, None, None, None


def table_gradient(table, rows_id, rows, sparse):
    """Row-sparse gradient of an embedding table from (row id, gradient row) pairs: duplicates are combined by a
    stable sort + ordered segmented sum (bit-reproducible).  sparse=True returns a torch sparse COO tensor whose
    padding entries (past the number of unique rows) are zero rows at index 0 (written by the kernel, no sync); sparse=False scatters into a dense
    zero tensor (the reference's dense `nn.Embedding` gradient)."""
    uid, urows, num = ops.sparse_rows_combine(rows_id, rows, table.shape[0])
    if not sparse:
        dense = torch.zeros_like(table)
        ops.scatter_rows(uid, urows, num, dense, accumulate=False)
        return dense
    return torch.sparse_coo_tensor(uid.unsqueeze(0), urows, table.shape, check_invariants=False)


class DirectEncoder(nn.Module):
    """Encodes a node as its (normalised) embedding-table row.

    features        -- callable (nodes, mode) -> rows, kept for signature compatibility; when it carries a
                       `node_maps` attribute (as `data_utils.build_graph` provides) or `node_maps` is passed, the
                       lookup runs in the fused CUDA kernel.
    feature_modules -- {mode: nn.Embedding}; registered as `feat-<mode>` like the reference (state_dict names)."""

    def __init__(self, features, feature_modules, node_maps=None, sparse_grad=False):
        super(DirectEncoder, self).__init__()
        for name, module in feature_modules.items():
            self.add_module('feat-' + name, module)
        self.features = features
        self.feature_modules = feature_modules
        if node_maps is None:
            node_maps = getattr(features, 'node_maps', None)
        self.register_buffer('node_maps', node_maps, persistent=False)
        self.sparse_grad = sparse_grad

    def table(self, mode):
        return self.feature_modules[mode].weight

    def ids_on_device(self, nodes, device):
        if torch.is_tensor(nodes):
            return nodes.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()
        return torch.as_tensor(nodes, dtype=torch.int64).to(device, non_blocking=True)

    def forward(self, nodes, mode, offset=None, **kwargs):
        if offset is not None:
            raise NotImplementedError('EmbeddingBag-style offsets are only used by the GQE encoders (out of scope)')
        table = self.table(mode)
        ops.device_guard(table.device)
        ids = self.ids_on_device(nodes, table.device)
        out = _GatherNormalize.apply(table, self.node_maps, ids.reshape(-1), self.sparse_grad)
        return out.t()
