"""ctypes binding of the C ABI declared in include/os2d_b200.h (libos2d_b200.so, built in-tree).

There is deliberately no fallback: if the shared library is missing or an entry point fails, the
caller gets an exception.  PyTorch is only used by the callers for device memory and streams.
"""
import ctypes
import functools
import os

_LIB_NAME = "libos2d_b200.so"
_HERE = os.path.dirname(os.path.abspath(__file__))

_c_int = ctypes.c_int
_c_float = ctypes.c_float
_c_void_p = ctypes.c_void_p
_c_ll = ctypes.c_longlong

MAX_PYRAMID_LEVELS = 12


class PyramidLevel(ctypes.Structure):
    """os2d_pyramid_level of include/os2d_b200.h"""
    _fields_ = [("loc", _c_void_p), ("score", _c_void_p), ("corners", _c_void_p), ("num_anchors", _c_int), ("fm_w", _c_int),
                ("img_w", _c_float), ("img_h", _c_float), ("scale_x", _c_float), ("scale_y", _c_float), ("same_scale", _c_int)]


# name -> (restype, argtypes); mirrors include/os2d_b200.h one to one
SIGNATURES = {
    "os2d_b200_abi_version": (_c_int, []),
    "os2d_b200_last_error": (ctypes.c_char_p, []),
    "os2d_b200_num_sms": (_c_int, []),
    "os2d_b200_launch_count": (ctypes.c_ulonglong, []),
    "os2d_pack_class_features": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p,
                                          _c_void_p]),
    "os2d_pack_class_features_ragged": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p,
                                                 _c_void_p]),
    "os2d_pack_image_features": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p]),
    "os2d_pack_image_features_nhwc": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_ll, _c_int, _c_void_p, _c_void_p]),
    "os2d_correlate": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p,
                                _c_void_p]),
    "os2d_correlate_conv1_concurrent": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p,
                                                 _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p,
                                                 _c_void_p]),
    "os2d_conv_weight_blob_bytes": (ctypes.c_size_t, [_c_int, _c_int]),
    "os2d_conv3_weight_blob_bytes": (ctypes.c_size_t, [_c_int]),
    "os2d_transform_conv": (_c_int, [_c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int,
                                     _c_int, _c_int, _c_void_p]),
    "os2d_resample_boxes": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float,
                                     _c_float, _c_float, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_ll, _c_ll,
                                     _c_void_p]),
    "os2d_resample_boxes_p2p": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float,
                                         _c_float, _c_float, _c_void_p, _c_int, _c_ll, _c_ll, _c_ll, _c_ll, _c_void_p]),
    "os2d_pack_corr_maps": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p]),
    "os2d_affine_grids": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p]),
    "os2d_resample_with_grid": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                                         _c_void_p]),
    "os2d_decode_boxes": (_c_int, [_c_int, _c_int, _c_int, _c_float, _c_float, _c_float, _c_float, _c_float, _c_float,
                                   _c_float, _c_float, _c_float, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                   _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "os2d_detect_pyramid": (_c_int, [ctypes.POINTER(PyramidLevel), _c_int, _c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_float,
                                     _c_float, _c_float, _c_float, _c_float, ctypes.c_double, _c_void_p, _c_void_p, _c_void_p,
                                     _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "os2d_gather_detections": (_c_int, [ctypes.POINTER(PyramidLevel), _c_int, _c_int, _c_void_p, _c_int, _c_float, _c_float,
                                        _c_float, _c_float, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                        _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "os2d_resize_level": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_int, _c_void_p,
                                   _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "os2d_voc_match": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_float,
                                _c_void_p, _c_void_p]),
    "os2d_nms_segments": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, ctypes.c_double, _c_void_p, _c_void_p]),
}


class Os2dB200Error(RuntimeError):
    pass


_lib = None


def library_path():
    return os.path.join(_HERE, _LIB_NAME)


def load():
    """Load (once) and return the ctypes handle; raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise Os2dB200Error(
            "{} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C os2d_b200/csrc` (no CPU or library fallback exists)".format(path))
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().os2d_b200_last_error()
        raise Os2dB200Error("{} failed with code {}: {}".format(what, rc, msg.decode() if msg else ""))


def ptr(t):
    """Device pointer of a (contiguous) torch tensor, or None.  Pass NAMED tensors: a temporary created inside the call
    expression is returned to the caching allocator before the launch and may be handed out to the next temporary."""
    if t is None:
        return None
    assert t.is_contiguous(), "os2d_b200 kernels take contiguous tensors"
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    """Current stream of the CURRENT device: entry points run under ``on_device_of`` so this is the tensors' device."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _first_cuda_tensor(obj, depth=0):
    import torch
    if isinstance(obj, torch.Tensor):
        return obj if obj.device.type == "cuda" else None
    if depth < 2 and isinstance(obj, (list, tuple)):
        for o in obj:
            t = _first_cuda_tensor(o, depth + 1)
            if t is not None:
                return t
    if hasattr(obj, "bbox_xyxy"):
        return _first_cuda_tensor(obj.bbox_xyxy, depth + 1)
    if hasattr(obj, "packed"):                       # head.PackedFeatureMaps
        return _first_cuda_tensor(obj.packed, depth + 1)
    return None


def on_device_of(fn):
    """Decorator of the public entry points: run ``fn`` with the CUDA device of its first CUDA tensor argument current.
    The C ABI launches on the current device's stream and caches per-device state keyed on the current device, so a call
    with tensors on cuda:1 while cuda:0 is current must switch first (and switch back afterwards)."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        import torch
        t = None
        for a in list(args) + list(kwargs.values()):
            t = _first_cuda_tensor(a)
            if t is not None:
                break
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(t.device):
            return fn(*args, **kwargs)
    return wrapper
