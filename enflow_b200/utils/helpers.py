"""Segment ops and small math of `enflow/utils/helpers.py`, on the CUDA path where it matters."""
import math

import torch

from .. import _lib


def log_gaussian(z):
    """`helpers.py:4-5` (log 2pi added once per call)."""
    return -0.5 * ((z ** 2).sum() + math.log(2 * math.pi))


def apply_pbc(pos, box):
    """`helpers.py:7-8`."""
    return pos - (pos / box).round() * box


def one_hot(index, num_classes=None, dtype=None):
    """`helpers.py:43-52`."""
    if index.dim() != 1:
        raise ValueError("'index' tensor needs to be one-dimensional")
    if num_classes is None:
        num_classes = int(index.max()) + 1
    out = torch.zeros((index.size(0), num_classes), dtype=dtype, device=index.device)
    return out.scatter_(1, index.unsqueeze(1), 1)


def _rowptr(segment_ids, num_segments):
    counts = torch.bincount(segment_ids, minlength=num_segments)
    ptr = torch.zeros(num_segments + 1, dtype=torch.int32, device=segment_ids.device)
    ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
    return ptr


def _segment(data, segment_ids, num_segments, mean):
    _lib.require_cuda(data, segment_ids)
    if segment_ids.numel() > 1 and not bool((segment_ids[1:] >= segment_ids[:-1]).all()):
        raise ValueError('enflow_b200 segment ops need row-sorted segment_ids (the CUDA neighbour list provides them)')
    L = _lib.lib()
    x = _lib.f32c(data)
    E, W = x.shape
    ptr = _rowptr(segment_ids.long(), num_segments)
    out = torch.empty(num_segments, W, dtype=torch.float32, device=x.device)
    if W == 128 and not mean:
        _lib.check(L.enflow_segment_sum128(_lib.ptr(x), _lib.ptr(ptr), None, num_segments, E, 0, _lib.ptr(out), _lib.stream()))
    elif W == 3:
        _lib.check(L.enflow_segment_sum3(_lib.ptr(x), _lib.ptr(ptr), None, num_segments, E, int(mean), 1.0, 0,
                                         _lib.ptr(out), _lib.stream()))
    else:
        raise ValueError(f'enflow_b200 segment ops support width 128 (sum) and 3 (sum/mean); got {W}')
    return out.to(data.dtype)


def unsorted_segment_sum(data, segment_ids, num_segments):
    """`helpers.py:54-60`, deterministic (fixed edge order) instead of scatter_add_ atomics."""
    return _segment(data, segment_ids, num_segments, mean=False)


def unsorted_segment_mean(data, segment_ids, num_segments):
    """`helpers.py:63-70`, count clamped to >= 1."""
    return _segment(data, segment_ids, num_segments, mean=True)
