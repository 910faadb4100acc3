"""Test-side stand-in for the slice of `torch_geometric` the reference imports
(`/root/reference/mpqe/model.py:201-203`, `/root/reference/mpqe/data_utils.py:9`).  TEST INFRASTRUCTURE ONLY."""
