"""`python -m yolov3 ...` / the `yolov3` console script (reference: setup.py:16-18 ->
yolov3/__main__.py:36-212) on the B200 hot path: same flags, batched image / video loops."""
from yolov3_b200.cli import main

if __name__ == "__main__":
    main()
