"""PyG `inits.uniform(size, tensor)`: U(-1/sqrt(size), 1/sqrt(size)); no-op for None
(call sites: /root/reference/mpqe/model.py:263-267)."""
import math


def uniform(size, tensor):
    if tensor is not None:
        bound = 1.0 / math.sqrt(size)
        tensor.data.uniform_(-bound, bound)
