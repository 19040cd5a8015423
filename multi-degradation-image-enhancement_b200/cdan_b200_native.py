"""ctypes binding of libcdan_b200.so (the C ABI declared in include/cdan_b200.h).

PyTorch is used only for device memory and streams: tensors are passed as raw ``data_ptr()`` values and the
current CUDA stream handle.  If the shared library is missing this module raises — there is no PyTorch / CPU
fallback behind it.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Iterable, Optional, Tuple

import torch

DTYPE_F32 = 0
DTYPE_BF16 = 1
_DTYPE_NAMES = {"fp32": DTYPE_F32, "float32": DTYPE_F32, "f32": DTYPE_F32, "bf16": DTYPE_BF16, "bfloat16": DTYPE_BF16}

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[ctypes.CDLL] = None

_c_int, _c_void_p, _c_char_p, _c_float = ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_float
_PROTOTYPES = {
    "cdan_last_error": (_c_char_p, []),
    "cdan_version": (_c_char_p, []),
    "cdan_plan_create": (_c_int, [_c_int, _c_int, ctypes.POINTER(_c_void_p)]),
    "cdan_plan_destroy": (_c_int, [_c_void_p]),
    "cdan_plan_load_weights": (_c_int, [_c_void_p, _c_int, ctypes.POINTER(_c_char_p), ctypes.POINTER(_c_void_p),
                                        ctypes.POINTER(ctypes.c_int64)]),
    "cdan_plan_set_option": (_c_int, [_c_void_p, _c_char_p, _c_int]),
    "cdan_workspace_bytes": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, ctypes.POINTER(ctypes.c_size_t)]),
    "cdan_forward": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int]),
    "cdan_forward_host": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int]),
    "cdan_forward_host_u8": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int]),
    "cdan_band_group_create": (_c_int, [_c_int, ctypes.POINTER(_c_void_p)]),
    "cdan_band_group_destroy": (_c_int, [_c_void_p]),
    "cdan_plan_band_attach_local": (_c_int, [_c_void_p, _c_void_p, _c_int]),
    "cdan_band_nccl_unique_id": (_c_int, [_c_void_p, ctypes.c_size_t]),
    "cdan_plan_band_attach_nccl": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, ctypes.c_size_t]),
    "cdan_plan_band_detach": (_c_int, [_c_void_p]),
    "cdan_band_rows": (_c_int, [_c_int, _c_int, _c_int, _c_int, ctypes.POINTER(_c_int)]),
    "cdan_forward_band": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int]),
    "cdan_band_stats": (_c_int, [_c_void_p, ctypes.POINTER(ctypes.c_longlong)]),
    "cdan_stage_read": (_c_int, [_c_void_p, _c_void_p, _c_char_p, _c_void_p, ctypes.POINTER(ctypes.c_int64)]),
    "cdan_last_launch_count": (_c_int, [_c_void_p]),
    "cdan_profile_read": (_c_int, [_c_void_p, ctypes.c_char_p, ctypes.c_size_t]),
    "cdan_op_conv2d": (_c_int, [_c_int, _c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                                _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p]),
    "cdan_op_cbam": (_c_int, [_c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p,
                              _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "cdan_op_upsample_add": (_c_int, [_c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int,
                                      _c_void_p]),
    "cdan_postprocess": (_c_int, [_c_void_p, _c_int, _c_float, _c_void_p, _c_void_p, _c_int, _c_int, _c_int]),
    "cdan_quantize_u8": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int]),
    "cdan_resize_normalize_u8": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_int, _c_int]),
    "cdan_psnr_ssim": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int,
                                ctypes.POINTER(_c_float)]),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)


def library_path() -> str:
    return os.environ.get("CDAN_B200_LIB", os.path.join(_HERE, "csrc", "libcdan_b200.so"))


def lib() -> ctypes.CDLL:
    """Load the shared library once; fail loudly when it is absent."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"cdan_b200: native library not found at {path}. Build it with "
                f"`bash {os.path.join(_HERE, 'csrc', 'build.sh')}` (nvcc, sm_100a). "
                "The CUDA kernels are the product: there is no PyTorch or CPU fallback.")
        handle = ctypes.CDLL(path)
        for name, (restype, argtypes) in _PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = restype, argtypes
        _LIB = handle
    return _LIB


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"cdan_b200::{what} failed: {lib().cdan_last_error().decode(errors='replace')}")


def dtype_code(dtype) -> int:
    if isinstance(dtype, int):
        return dtype
    if isinstance(dtype, torch.dtype):
        return {torch.float32: DTYPE_F32, torch.bfloat16: DTYPE_BF16}[dtype]
    return _DTYPE_NAMES[str(dtype).lower()]


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _f32(t: torch.Tensor, device: torch.device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class Plan:
    """One (device, dtype) execution plan: packed weights + workspace inside the native library."""

    def __init__(self, device: torch.device, dtype="bf16"):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("cdan_b200: plans exist only on CUDA devices (no CPU fallback)")
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        self.dtype = dtype_code(dtype)
        self._h = ctypes.c_void_p()
        _check(lib().cdan_plan_create(self.device.index, self.dtype, ctypes.byref(self._h)), "plan_create")

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            lib().cdan_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name: str, value: int) -> None:
        _check(lib().cdan_plan_set_option(self._h, name.encode(), int(value)), "plan_set_option")

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor]) -> None:
        items = [(k, v) for k, v in state_dict.items() if not k.endswith("num_batches_tracked")]
        keep = [v.detach().to(torch.float32).contiguous() for _, v in items]  # host or device, both accepted
        n = len(items)
        keys = (_c_char_p * n)(*[k.encode() for k, _ in items])
        ptrs = (_c_void_p * n)(*[t.data_ptr() for t in keep])
        numels = (ctypes.c_int64 * n)(*[t.numel() for t in keep])
        if any(t.is_cuda for t in keep):
            torch.cuda.synchronize(self.device)
        _check(lib().cdan_plan_load_weights(self._h, n, keys, ptrs, numels), "plan_load_weights")

    def workspace_bytes(self, n: int, h: int, w: int) -> int:
        out = ctypes.c_size_t()
        _check(lib().cdan_workspace_bytes(self._h, n, h, w, ctypes.byref(out)), "workspace_bytes")
        return out.value

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"cdan_b200: expected input [N,3,H,W], got {tuple(x.shape)}")
        if x.device != self.device:
            raise RuntimeError(f"cdan_b200: input on {x.device}, plan on {self.device}")
        x = x.detach().to(torch.float32).contiguous()
        y = out if out is not None else torch.empty_like(x)
        n, _, h, w = x.shape
        with torch.cuda.device(self.device):
            _check(lib().cdan_forward(self._h, _c_void_p(_stream(self.device)), _ptr(x), _ptr(y), n, h, w), "forward")
        return y

    def forward_host(self, x_host: torch.Tensor, out_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host buffers in/out (pinned memory recommended); copies + forward + sync happen in the library."""
        if x_host.is_cuda or x_host.dtype != torch.float32 or not x_host.is_contiguous():
            raise RuntimeError("cdan_b200: forward_host needs a contiguous fp32 CPU tensor")
        y = out_host if out_host is not None else torch.empty_like(x_host)
        n, _, h, w = x_host.shape
        _check(lib().cdan_forward_host(self._h, _ptr(x_host), _ptr(y), n, h, w), "forward_host")
        return y

    def forward_host_u8(self, x_host: torch.Tensor, out_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 [N,H,W,3] host buffers in/out: the reference's image data path (uint8 -> /255 -> forward -> x255 ->
        uint8) with normalisation and quantisation on the device; a quarter of the PCIe bytes of forward_host."""
        if x_host.is_cuda or x_host.dtype != torch.uint8 or not x_host.is_contiguous() or x_host.dim() != 4 or x_host.shape[3] != 3:
            raise RuntimeError("cdan_b200: forward_host_u8 needs a contiguous uint8 CPU tensor [N,H,W,3]")
        y = out_host if out_host is not None else torch.empty_like(x_host)
        n, h, w, _ = x_host.shape
        _check(lib().cdan_forward_host_u8(self._h, _ptr(x_host), _ptr(y), n, h, w), "forward_host_u8")
        return y

    # ---- spatial row tiling (include/cdan_b200.h "spatial row tiling"; host side in spatial_tiling.py)
    def band_attach_local(self, group: "BandGroup", rank: int) -> None:
        _check(lib().cdan_plan_band_attach_local(self._h, group._h, int(rank)), "plan_band_attach_local")
        self._band_group = group  # keep the transport alive as long as the plan uses it

    def band_attach_nccl(self, rank: int, nranks: int, unique_id: bytes) -> None:
        buf = ctypes.create_string_buffer(bytes(unique_id), len(unique_id))
        with torch.cuda.device(self.device):
            _check(lib().cdan_plan_band_attach_nccl(self._h, int(rank), int(nranks), buf, len(unique_id)), "plan_band_attach_nccl")

    def band_detach(self) -> None:
        lib().cdan_plan_band_detach(self._h)
        self._band_group = None

    def forward_band(self, x_ext: torch.Tensor, height: int, halo: int, out: Optional[torch.Tensor] = None,
                     stream: Optional[int] = None) -> torch.Tensor:
        """x_ext: rows [extended begin, extended end) of the [N,3,height,W] image (band_rows); collective over the bands."""
        if x_ext.dim() != 4 or x_ext.shape[1] != 3 or x_ext.device != self.device or x_ext.dtype != torch.float32 or not x_ext.is_contiguous():
            raise RuntimeError("cdan_b200: forward_band needs a contiguous fp32 [N,3,Hext,W] tensor on the plan's device")
        y = out if out is not None else torch.empty_like(x_ext)
        n, _, _, w = x_ext.shape
        st = _stream(self.device) if stream is None else stream
        _check(lib().cdan_forward_band(self._h, _c_void_p(st), _ptr(x_ext), _ptr(y), n, int(height), w, int(halo)), "forward_band")
        return y

    def band_stats(self) -> Dict[str, int]:
        out = (ctypes.c_longlong * 3)()
        _check(lib().cdan_band_stats(self._h, out), "band_stats")
        return {"halo_exchanges": int(out[0]), "halo_bytes_received": int(out[1]), "allreduces": int(out[2])}

    def stage(self, name: str) -> torch.Tensor:
        shape = (ctypes.c_int64 * 4)()
        s = _c_void_p(_stream(self.device))
        _check(lib().cdan_stage_read(self._h, s, name.encode(), None, shape), "stage_read")
        out = torch.empty(tuple(shape), dtype=torch.float32, device=self.device)
        _check(lib().cdan_stage_read(self._h, s, name.encode(), _ptr(out), shape), "stage_read")
        return out

    def profile_read(self) -> Dict[str, Tuple[float, int]]:
        """Drain the per-launch CUDA-event spans recorded since the last call (needs set_option('profile', 1))."""
        buf = ctypes.create_string_buffer(1 << 16)
        _check(lib().cdan_profile_read(self._h, buf, len(buf)), "profile_read")
        out: Dict[str, Tuple[float, int]] = {}
        for line in buf.value.decode().splitlines():
            label, ms, cnt = line.rsplit(" ", 2)
            out[label] = (float(ms), int(cnt))
        return out

    @property
    def last_launch_count(self) -> int:
        return lib().cdan_last_launch_count(self._h)


class BandGroup:
    """In-process transport between the bands of one process (one host thread per band drives its plan)."""

    def __init__(self, nbands: int):
        self._h = ctypes.c_void_p()
        self.nbands = int(nbands)
        _check(lib().cdan_band_group_create(self.nbands, ctypes.byref(self._h)), "band_group_create")

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            lib().cdan_band_group_destroy(self._h)
            self._h = ctypes.c_void_p()


def band_rows(height: int, nbands: int, rank: int, halo: int) -> Tuple[int, int, int, int]:
    """(owned begin, owned end, extended begin, extended end) — the library's band split (cdan_band_rows)."""
    out = (_c_int * 4)()
    _check(lib().cdan_band_rows(int(height), int(nbands), int(rank), int(halo), out), "band_rows")
    return tuple(int(v) for v in out)


def nccl_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    _check(lib().cdan_band_nccl_unique_id(buf, 128), "band_nccl_unique_id")
    return buf.raw


# ------------------------------------------------------------------------------------------------ single operators
def op_conv2d(x, w, bias=None, pre_scale=None, pre_shift=None, relu=False, pool=False, dtype="fp32", impl=0):
    dev = x.device
    x, w = _f32(x, dev), _f32(w, dev)
    bias = None if bias is None else _f32(bias, dev)
    pre_scale = None if pre_scale is None else _f32(pre_scale, dev)
    pre_shift = None if pre_shift is None else _f32(pre_shift, dev)
    n, cin, h, wd = x.shape
    cout, _, ks, _ = w.shape
    y = torch.empty((n, cout, h // 2 if pool else h, wd // 2 if pool else wd), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _check(lib().cdan_op_conv2d(dtype_code(dtype), impl, _c_void_p(_stream(dev)), _ptr(x), n, cin, h, wd, _ptr(w),
                                    _ptr(bias), cout, ks, _ptr(pre_scale), _ptr(pre_shift), int(relu), int(pool),
                                    _ptr(y)), "op_conv2d")
    return y


def op_cbam(x, w1, b1, w2, b2, w7, bn4: Iterable[float], mul=None, dtype="fp32"):
    dev = x.device
    x = _f32(x, dev)
    ws = [_f32(t, dev) for t in (w1, b1, w2, b2, w7)]
    mul = None if mul is None else _f32(mul, dev)
    bn = (ctypes.c_float * 4)(*[float(v) for v in bn4])
    n, c, h, w = x.shape
    y = torch.empty_like(x)
    with torch.cuda.device(dev):
        _check(lib().cdan_op_cbam(dtype_code(dtype), _c_void_p(_stream(dev)), _ptr(x), n, c, h, w, *[_ptr(t) for t in ws],
                                  ctypes.cast(bn, _c_void_p), _ptr(mul), _ptr(y)), "op_cbam")
    return y


def op_upsample_add(a, skip, up=True, dtype="fp32"):
    dev = a.device
    a, skip = _f32(a, dev), _f32(skip, dev)
    n, c, h, w = a.shape
    y = torch.empty_like(skip)
    with torch.cuda.device(dev):
        _check(lib().cdan_op_upsample_add(dtype_code(dtype), _c_void_p(_stream(dev)), _ptr(a), _ptr(skip), n, c, h, w,
                                          int(up), _ptr(y)), "op_upsample_add")
    return y


POSTPROC_OPS = {"enhance_contrast": 0, "enhance_color": 1, "sharpen": 2, "soft_denoise": 3}


def postprocess(images: torch.Tensor, op: str, arg: float) -> torch.Tensor:
    dev = images.device
    x = _f32(images, dev)
    if x.dim() != 4 or x.shape[1] != 3:
        raise RuntimeError("cdan_b200: post-processing expects [N,3,H,W]")
    y = torch.empty_like(x)
    n, _, h, w = x.shape
    with torch.cuda.device(dev):
        _check(lib().cdan_postprocess(_c_void_p(_stream(dev)), POSTPROC_OPS[op], float(arg), _ptr(x), _ptr(y), n, h, w),
               "postprocess")
    return y


def quantize_u8(images: torch.Tensor) -> torch.Tensor:
    """[N,3,H,W] float (CUDA) -> [N,H,W,3] uint8 (CUDA): `(img * 255).clip(0, 255).astype(uint8)` of the reference's
    _save_batch_outputs (models/model.py:80-83), on the device, so that only a quarter of the bytes cross PCIe."""
    dev = images.device
    x = _f32(images, dev)
    if x.dim() != 4 or x.shape[1] != 3:
        raise RuntimeError("cdan_b200: quantize_u8 expects [N,3,H,W]")
    n, _, h, w = x.shape
    y = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _check(lib().cdan_quantize_u8(_c_void_p(_stream(dev)), _ptr(x), _ptr(y), n, h, w), "quantize_u8")
    return y


def resize_normalize_u8(images_u8: torch.Tensor, out_hw) -> torch.Tensor:
    """[N,Hs,Ws,3] uint8 (CUDA) -> [N,3,Hd,Wd] float32 (CUDA): the reference's input transform `A.Resize(Hd, Wd)`
    (cv2.resize INTER_LINEAR on the uint8 image, bit-exact) + `A.Normalize(0, 1, 255)` + `ToTensorV2`
    (utils/transforms_factory.py:50-86), on the device — the uint8 source is a quarter of the bytes over PCIe."""
    if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[3] != 3 or not images_u8.is_cuda:
        raise RuntimeError("cdan_b200: resize_normalize_u8 expects a CUDA uint8 tensor [N,H,W,3]")
    x = images_u8.contiguous()
    n, hs, ws, _ = x.shape
    hd, wd = int(out_hw[0]), int(out_hw[1])
    y = torch.empty((n, 3, hd, wd), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _check(lib().cdan_resize_normalize_u8(_c_void_p(_stream(x.device)), _ptr(x), n, hs, ws, _ptr(y), hd, wd),
               "resize_normalize_u8")
    return y


def psnr_ssim(pred: torch.Tensor, target: torch.Tensor) -> Tuple[float, float]:
    dev = pred.device
    p, t = _f32(pred, dev), _f32(target, dev)
    n, c, h, w = p.shape
    res = (ctypes.c_float * 2)()
    with torch.cuda.device(dev):
        _check(lib().cdan_psnr_ssim(_c_void_p(_stream(dev)), _ptr(p), _ptr(t), n, c, h, w, res), "psnr_ssim")
    return float(res[0]), float(res[1])
