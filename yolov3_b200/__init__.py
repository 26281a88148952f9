"""yolov3_b200 — Blackwell-native (sm_100a) YOLOv3 detection hot path behind the ultralytics/yolov3 Python seams.

Public surface mirrors the reference's: ``Model`` (models/yolo.py:193), ``non_max_suppression``
(utils/general.py:630), ``ComputeLoss`` (utils/loss.py:98), ``box_iou`` (utils/metrics.py:10).  All compute runs in the
hand-written CUDA library ``libyolov3_b200.so`` through the C ABI in ``include/yolov3_b200.h``.
"""
__version__ = "0.1.0"
