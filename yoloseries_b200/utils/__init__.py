"""Mirror of the reference's ``utils`` names that sit on the post-processing hot path (utils/__init__.py:1-21)."""
from .bbox_tools import gpu_CIoU, gpu_DIoU, gpu_Giou, gpu_iou, numba_iou  # noqa: F401
from . import mAP  # noqa: F401
from .mAP import compute_tp, compute_tp_batch  # noqa: F401
from .weighted_fusion_bbox import weighted_fusion_bbox  # noqa: F401
from .nms import gpu_exponential_soft_nms, gpu_linear_soft_nms, gpu_nms, numba_nms  # noqa: F401

__all__ = ["numba_nms", "gpu_nms", "gpu_linear_soft_nms", "gpu_exponential_soft_nms", "numba_iou", "gpu_iou", "gpu_Giou", "gpu_DIoU", "gpu_CIoU", "compute_tp", "compute_tp_batch", "weighted_fusion_bbox"]
