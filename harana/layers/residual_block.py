from svcc23_fastsvc_b200.layers import Conv1d, Conv1d1x1  # noqa: F401

__all__ = ["Conv1d", "Conv1d1x1"]
