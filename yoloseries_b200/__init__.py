"""yoloseries_b200 -- B200-native (sm_100a) detection post-processing for yl-jiang/YOLOSeries' hot path.

decode raw heads -> filter by conf x cls -> top-k -> pairwise IoU -> class-aware NMS -> [K, 6] rows per image, behind
the reference's own Python signatures (``yoloseries_b200.utils`` / ``yoloseries_b200.trainer``) over the C ABI of
``include/ysb_postproc.h``.  There is no CPU or PyTorch fallback: importing works anywhere, computing needs the built
library (``python -m yoloseries_b200.build``) and a CUDA device.
"""
__version__ = "0.1.0"
