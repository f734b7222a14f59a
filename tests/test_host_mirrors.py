"""Argument handling of the reference-facing mirrors that is decided on the host, before any device work (no GPU):
the same exceptions / trivial results the reference gives (utils/nms.py:12,38-39,53; SURVEY.md 8b conventions)."""
import numpy as np
import pytest
import torch

from yoloseries_b200.utils import (gpu_CIoU, gpu_Giou, gpu_exponential_soft_nms, gpu_linear_soft_nms, gpu_nms, numba_nms)


def test_length_mismatch_asserts_like_the_reference():
    with pytest.raises(AssertionError):
        numba_nms(np.zeros((3, 4), np.float32), np.zeros(2, np.float32), 0.5)
    with pytest.raises(AssertionError):
        gpu_nms(torch.zeros(3, 4), torch.zeros(2), "giou", 0.5)
    with pytest.raises(AssertionError):
        gpu_nms(np.zeros((3, 4), np.float32), torch.zeros(3), "giou", 0.5)     # tensors only (utils/nms.py:38)
    with pytest.raises(AssertionError):
        gpu_linear_soft_nms(torch.zeros(3, 4), torch.zeros(2, 1), "giou")


def test_unknown_iou_type_is_a_value_error():
    with pytest.raises(ValueError):
        gpu_nms(torch.zeros(3, 4), torch.zeros(3), "siou", 0.5)
    with pytest.raises(ValueError):
        gpu_exponential_soft_nms(torch.zeros(3, 4), torch.zeros(3, 1), "siou", 0.3)


def test_empty_inputs_need_no_device():
    assert numba_nms(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 0.5) == []
    assert gpu_nms(torch.zeros(0, 4), torch.zeros(0), "diou", 0.5) == []


def test_rowwise_iou_shape_checks():
    with pytest.raises(AssertionError):
        gpu_CIoU(torch.zeros(3, 5), torch.zeros(3, 5))                           # last dimension must be 4
    with pytest.raises(AssertionError):
        gpu_Giou(torch.tensor([[2.0, 0.0, 1.0, 1.0]]), torch.tensor([[0.0, 0.0, 1.0, 1.0]]))   # x2 >= x1 (:201)
    with pytest.raises(RuntimeError):
        gpu_CIoU(torch.zeros(2, 4), torch.zeros(3, 4))                           # neither equal nor broadcastable


def test_compute_needs_a_device():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gpu_nms(torch.rand(4, 4), torch.rand(4), "giou", 0.5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gpu_linear_soft_nms(torch.rand(4, 4), torch.rand(4, 1), "giou")
