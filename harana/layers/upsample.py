from svcc23_fastsvc_b200.layers import Conv1d1x3, Conv2d1x3, Squeeze2d, Stretch2d  # noqa: F401

__all__ = ["Stretch2d", "Squeeze2d", "Conv1d1x3", "Conv2d1x3"]
