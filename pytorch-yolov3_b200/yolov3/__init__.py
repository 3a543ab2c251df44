"""Drop-in alias: ``import yolov3`` resolves to the B200-native hot path when
``pytorch-yolov3_b200/`` precedes the reference on ``sys.path``;
``python -m yolov3 -I images/ -c cfg -w weights`` is the batched CLI (``yolov3_b200/cli.py``).
The marker file ``_b200_alias`` lets ``cli.reference_package()`` tell this alias from the reference."""
from yolov3_b200 import *  # noqa: F401,F403
from yolov3_b200 import __all__, darknet, inference as _inference_fn  # noqa: F401
