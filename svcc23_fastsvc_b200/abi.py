"""ctypes binding of libfsvc.so (C ABI declared in include/fsvc.h).

The library is built in-tree by ``svcc23_fastsvc_b200.build`` (nvcc, sm_100a).
There is no CPU fallback: a missing library or a missing CUDA device raises.
"""

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FSVC_LIB") or os.path.join(_HERE, "libfsvc.so")   # FSVC_LIB: a debug build (timeline)
MAX_STAGES = 8

MODE_FP32 = 0
MODE_TC_BF16X3 = 1
MODE_AUTO = 2
MODES = {"fp32": MODE_FP32, "tc_bf16x3": MODE_TC_BF16X3, "auto": MODE_AUTO}

# every symbol include/fsvc.h declares
EXPORTS = (
    "fsvc_abi_version", "fsvc_last_error", "fsvc_create", "fsvc_destroy", "fsvc_num_weight_tensors",
    "fsvc_weight_tensor_info", "fsvc_set_weights", "fsvc_workspace_bytes", "fsvc_forward",
    "fsvc_host_io_bytes", "fsvc_forward_host", "fsvc_downsample_forward", "fsvc_film_forward",
    "fsvc_upsample_forward", "fsvc_block_workspace_bytes", "fsvc_last_launch_count", "fsvc_forward_profile",
    "fsvc_sine_excitation", "fsvc_pcm16", "fsvc_train_saved_bytes", "fsvc_train_workspace_bytes",
    "fsvc_forward_train", "fsvc_backward",
)


class KernelRecord(ctypes.Structure):
    _fields_ = [("label", ctypes.c_char * 48), ("ms", ctypes.c_float), ("flops", ctypes.c_double),
                ("bytes", ctypes.c_double)]


class FsvcConfig(ctypes.Structure):
    _fields_ = [
        ("in_channels", ctypes.c_int32),
        ("num_stages", ctypes.c_int32),
        ("mid_channels", ctypes.c_int32 * MAX_STAGES),
        ("upsampling_scales", ctypes.c_int32 * MAX_STAGES),
        ("out_channels", ctypes.c_int32),
        ("spk_emb_size", ctypes.c_int32),
        ("use_spk_emb", ctypes.c_int32),
        ("lrelu_slope", ctypes.c_float),
        ("in_eps", ctypes.c_float),
    ]


class FsvcError(RuntimeError):
    pass


_lib = None


def load():
    """Load libfsvc.so (once) and declare the prototypes of include/fsvc.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FsvcError(
            f"{LIB_PATH} not found: build it with `python -m svcc23_fastsvc_b200.build` "
            "(nvcc, sm_100a). The FastSVC generator has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, sz, fp = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float
    pp = ctypes.POINTER(ctypes.c_void_p)
    lib.fsvc_abi_version.restype = i32
    lib.fsvc_abi_version.argtypes = []
    lib.fsvc_last_error.restype = ctypes.c_char_p
    lib.fsvc_last_error.argtypes = []
    lib.fsvc_create.restype = i32
    lib.fsvc_create.argtypes = [ctypes.POINTER(FsvcConfig), ctypes.POINTER(vp)]
    lib.fsvc_destroy.restype = None
    lib.fsvc_destroy.argtypes = [vp]
    lib.fsvc_num_weight_tensors.restype = i32
    lib.fsvc_num_weight_tensors.argtypes = [vp]
    lib.fsvc_weight_tensor_info.restype = i32
    lib.fsvc_weight_tensor_info.argtypes = [vp, i32, ctypes.c_char_p, i32, ctypes.POINTER(ctypes.c_int64)]
    lib.fsvc_set_weights.restype = i32
    lib.fsvc_set_weights.argtypes = [vp, pp, i32, vp]
    lib.fsvc_workspace_bytes.restype = sz
    lib.fsvc_workspace_bytes.argtypes = [vp, i32, i32, i32]
    lib.fsvc_forward.restype = i32
    lib.fsvc_forward.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp, sz, i32, vp]
    lib.fsvc_host_io_bytes.restype = sz
    lib.fsvc_host_io_bytes.argtypes = [vp, i32, i32]
    lib.fsvc_forward_host.restype = i32
    lib.fsvc_forward_host.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp, sz, i32, vp]
    lib.fsvc_downsample_forward.restype = i32
    lib.fsvc_downsample_forward.argtypes = [vp, vp, pp, i32, i32, i32, i32, i32, fp, vp, sz, i32, vp]
    lib.fsvc_film_forward.restype = i32
    lib.fsvc_film_forward.argtypes = [vp, vp, vp, pp, i32, i32, i32, fp, vp, sz, i32, vp]
    lib.fsvc_upsample_forward.restype = i32
    lib.fsvc_upsample_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, pp, i32, i32, i32, i32, i32, i32, fp, fp,
                                          vp, sz, i32, vp]
    lib.fsvc_block_workspace_bytes.restype = sz
    lib.fsvc_block_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.fsvc_forward_profile.restype = i32
    lib.fsvc_forward_profile.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp, sz, i32, vp,
                                         ctypes.POINTER(KernelRecord), i32, ctypes.POINTER(i32)]
    lib.fsvc_sine_excitation.restype = i32
    lib.fsvc_sine_excitation.argtypes = [vp, vp, vp, i32, i32, i32, fp, fp, fp, vp]
    lib.fsvc_pcm16.restype = i32
    lib.fsvc_pcm16.argtypes = [vp, vp, ctypes.c_longlong, vp]
    lib.fsvc_last_launch_count.restype = i32
    lib.fsvc_last_launch_count.argtypes = [vp]
    lib.fsvc_train_saved_bytes.restype = sz
    lib.fsvc_train_saved_bytes.argtypes = [vp, i32, i32]
    lib.fsvc_train_workspace_bytes.restype = sz
    lib.fsvc_train_workspace_bytes.argtypes = [vp, i32, i32]
    lib.fsvc_forward_train.restype = i32
    lib.fsvc_forward_train.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp, sz, vp, sz, vp]
    lib.fsvc_backward.restype = i32
    lib.fsvc_backward.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp, sz, pp, i32, vp, sz, vp]
    if lib.fsvc_abi_version() != 1:
        raise FsvcError(f"libfsvc ABI version {lib.fsvc_abi_version()} != 1")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise FsvcError(f"libfsvc error {rc}: {load().fsvc_last_error().decode()}")


def ptr_array(ptrs):
    arr = (ctypes.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


class Handle:
    """Owns one ``fsvc_handle`` (one generator configuration on one device)."""

    def __init__(self, in_channels, mid_channels, upsampling_scales, out_channels, spk_emb_size, use_spk_emb,
                 lrelu_slope=0.2, in_eps=1e-5):
        lib = load()
        if len(mid_channels) != len(upsampling_scales):
            raise ValueError("mid_channels and upsampling_scales must have the same length")
        if len(mid_channels) > MAX_STAGES:
            raise ValueError(f"at most {MAX_STAGES} stages")
        cfg = FsvcConfig()
        cfg.in_channels = in_channels
        cfg.num_stages = len(mid_channels)
        for i, (c, r) in enumerate(zip(mid_channels, upsampling_scales)):
            cfg.mid_channels[i] = c
            cfg.upsampling_scales[i] = r
        cfg.out_channels = out_channels
        cfg.spk_emb_size = spk_emb_size
        cfg.use_spk_emb = 1 if use_spk_emb else 0
        cfg.lrelu_slope = lrelu_slope
        cfg.in_eps = in_eps
        self._h = ctypes.c_void_p()
        self._lib = lib
        check(lib.fsvc_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        n = lib.fsvc_num_weight_tensors(self._h)
        self.weight_names, self.weight_numel = [], []
        buf = ctypes.create_string_buffer(256)
        numel = ctypes.c_int64()
        for i in range(n):
            check(lib.fsvc_weight_tensor_info(self._h, i, buf, 256, ctypes.byref(numel)))
            self.weight_names.append(buf.value.decode())
            self.weight_numel.append(numel.value)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.fsvc_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weights(self, ptrs, stream):
        check(self._lib.fsvc_set_weights(self._h, ptr_array(ptrs), len(ptrs), stream))

    def workspace_bytes(self, B, frames, mode):
        n = self._lib.fsvc_workspace_bytes(self._h, B, frames, mode)
        if n == 0:
            raise FsvcError(f"libfsvc: {self._lib.fsvc_last_error().decode()}")
        return n

    def host_io_bytes(self, B, frames):
        return self._lib.fsvc_host_io_bytes(self._h, B, frames)

    def forward(self, ppg, sine, lft, spk, out, B, frames, ws, ws_bytes, mode, stream):
        check(self._lib.fsvc_forward(self._h, ppg, sine, lft, spk, out, B, frames, ws, ws_bytes, mode, stream))

    def forward_host(self, ppg, sine, lft, spk, out, B, frames, ws, ws_bytes, mode, stream):
        check(self._lib.fsvc_forward_host(self._h, ppg, sine, lft, spk, out, B, frames, ws, ws_bytes, mode, stream))

    def forward_profile(self, ppg, sine, lft, spk, out, B, frames, ws, ws_bytes, mode, stream, capacity=512):
        recs = (KernelRecord * capacity)()
        n = ctypes.c_int(0)
        check(self._lib.fsvc_forward_profile(self._h, ppg, sine, lft, spk, out, B, frames, ws, ws_bytes, mode,
                                             stream, recs, capacity, ctypes.byref(n)))
        return [dict(label=recs[i].label.decode(), ms=recs[i].ms, flops=recs[i].flops, bytes=recs[i].bytes)
                for i in range(n.value)]

    def last_launch_count(self):
        return self._lib.fsvc_last_launch_count(self._h)

    def train_saved_bytes(self, B, frames):
        return self._lib.fsvc_train_saved_bytes(self._h, B, frames)

    def train_workspace_bytes(self, B, frames):
        return self._lib.fsvc_train_workspace_bytes(self._h, B, frames)

    def forward_train(self, ppg, sine, lft, spk, out, B, frames, saved, saved_bytes, ws, ws_bytes, stream):
        check(self._lib.fsvc_forward_train(self._h, ppg, sine, lft, spk, out, B, frames, saved, saved_bytes, ws,
                                           ws_bytes, stream))

    def backward(self, ppg, sine, lft, spk, grad_out, B, frames, saved, saved_bytes, grad_ptrs, ws, ws_bytes, stream):
        check(self._lib.fsvc_backward(self._h, ppg, sine, lft, spk, grad_out, B, frames, saved, saved_bytes,
                                      ptr_array(grad_ptrs), len(grad_ptrs), ws, ws_bytes, stream))
