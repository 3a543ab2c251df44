"""Drop-in alias: ``import yolov3`` resolves to the B200-native hot path when
``pytorch-yolov3_b200/`` precedes the reference on ``sys.path``."""
from yolov3_b200 import *  # noqa: F401,F403
from yolov3_b200 import __all__, darknet, inference as _inference_fn  # noqa: F401
