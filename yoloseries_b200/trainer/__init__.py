"""Mirror of the reference's ``trainer`` exports on the post-processing path (trainer/__init__.py:1-8)."""
from .evaluators import (  # noqa: F401
    FCOSEvaluator,
    RetinaNetEvaluator,
    RetinaNetEvaluatorExperiment,
    YOLOV5Evaluator,
    YOLOV7Evaluator,
    YOLOV8Evaluator,
    YOLOXEvaluator,
)
