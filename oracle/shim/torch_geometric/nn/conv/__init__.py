"""PyG ~1.4 `MessagePassing` restated for the one way the reference drives it
(`/root/reference/mpqe/model.py:206, 236, 277`): aggr='add', flow source_to_target,
`propagate(edge_index, **kwargs)` -> `message(*)` with `<name>_j = kwargs[name][edge_index[0]]`,
`<name>_i = kwargs[name][edge_index[1]]`, other kwargs by name; scatter-add of the messages at
`edge_index[1]` with `dim_size = x.size(0)`; then `update(aggr_out, *)` with kwargs by name."""
import inspect

import torch


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr='add', flow='source_to_target'):
        super().__init__()
        assert aggr == 'add' and flow == 'source_to_target'
        self._msg_args = list(inspect.signature(self.message).parameters)
        self._upd_args = list(inspect.signature(self.update).parameters)[1:]

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        num_nodes = None
        margs = []
        for name in self._msg_args:
            if name.endswith('_j') or name.endswith('_i'):
                full = kwargs[name[:-2]]
                if full is not None and num_nodes is None:
                    num_nodes = full.size(0)
                sel = src if name.endswith('_j') else dst
                margs.append(None if full is None else full.index_select(0, sel))
            else:
                margs.append(kwargs[name])
        msg = self.message(*margs)
        if num_nodes is None:
            num_nodes = int(edge_index.max().item()) + 1
        out = torch.zeros((num_nodes,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        out = out.index_add(0, dst, msg)
        return self.update(out, *[kwargs[name] for name in self._upd_args])

    def message(self, x_j):  # pragma: no cover - overridden
        return x_j

    def update(self, aggr_out):  # pragma: no cover - overridden
        return aggr_out
