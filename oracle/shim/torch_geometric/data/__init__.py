"""PyG `Data` / `Batch.from_data_list` restated for `/root/reference/mpqe/data_utils.py:402-405`:
edge_index of graph g is shifted by the cumulative node count, every other tensor attribute is
concatenated, `batch[node] = g`."""
import torch


class Data(object):
    def __init__(self, x=None, edge_index=None, **kwargs):
        self.x = x
        self.edge_index = edge_index
        self.num_nodes = None
        for k, v in kwargs.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class Batch(Data):
    @staticmethod
    def from_data_list(data_list):
        out = Batch()
        shift = 0
        edge_index, batch, extra = [], [], {}
        for g, data in enumerate(data_list):
            n = data.num_nodes
            edge_index.append(data.edge_index + shift)
            batch.append(torch.full((n,), g, dtype=torch.long))
            for k, v in data.__dict__.items():
                if k in ('x', 'edge_index', 'num_nodes') or not torch.is_tensor(v):
                    continue
                extra.setdefault(k, []).append(v)
            shift += n
        out.edge_index = torch.cat(edge_index, dim=1)
        out.batch = torch.cat(batch)
        out.num_nodes = shift
        out.num_graphs = len(data_list)
        for k, vs in extra.items():
            setattr(out, k, torch.cat(vs, dim=0))
        return out
