"""Python view of the whole-network C entry points (include/b2r.h section 7): b2r_net_create + b2r_*_forward.

This is the path a NON-Python host uses (examples/cabi_pipeline.c): the layer graph and the weight packing live in
csrc/net_plan.cu.  The nn.Module classes in models.py issue the same launches layer by layer from Python (which is what
lets bench.py time every launch); tests/test_net_plan_gpu.py requires the two paths to agree bit for bit.
PyTorch is used here only to own the device buffers."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib as L

_ARCH = {"simple_unet": L.B2R_NET_SIMPLE_UNET, "resunet": L.B2R_NET_RESUNET, "vgg16": L.B2R_NET_VGG16}


def state_to_ctypes(state: Dict[str, torch.Tensor]):
    """(b2r_tensor array, keep-alive list) for a reference state_dict; tensors are taken as contiguous CPU float32 / int64."""
    keep, arr = [], (L.Tensor * len(state))()
    for i, (k, v) in enumerate(state.items()):
        t = v.detach().cpu().contiguous()
        if t.dtype not in (torch.float32, torch.int64):
            t = t.float()
        name = k.encode()
        keep += [t, name]
        arr[i].name = name
        arr[i].data = t.data_ptr()
        arr[i].dtype = L.B2R_DT_F32 if t.dtype == torch.float32 else L.B2R_DT_I64
        arr[i].ndim = t.dim()
        for d, s in enumerate(t.shape):
            arr[i].shape[d] = int(s)
    return arr, keep


class NetPlan:
    """b2r_net handle + the caller-owned device buffers it points into."""

    def __init__(self, arch: str, state: Dict[str, torch.Tensor], device, num_classes: int = 43):
        self.arch, self.device = arch, torch.device(device)
        lib = L.load()
        arr, keep = state_to_ctypes(state)
        nbytes = C.c_size_t()
        L.check(lib.b2r_net_weight_bytes(arr, len(state), C.byref(nbytes)))
        with torch.cuda.device(self.device):
            self.weights = torch.empty(int(nbytes.value), dtype=torch.uint8, device=self.device)
            handle = C.c_void_p()
            L.check(lib.b2r_net_create(_ARCH[arch], int(num_classes), arr, len(state), self.weights.data_ptr(), int(nbytes.value),
                                       torch.cuda.current_stream().cuda_stream, C.byref(handle)))
        self.handle, self.num_classes = handle, int(num_classes)
        self._ws: Optional[torch.Tensor] = None
        del keep

    def close(self):
        if getattr(self, "handle", None):
            L.load().b2r_net_destroy(self.handle)
            self.handle = None

    __del__ = close

    def _workspace(self, n, h, w):
        need = C.c_size_t()
        L.check(L.load().b2r_net_workspace_bytes(self.handle, n, h, w, C.byref(need)))
        if self._ws is None or self._ws.numel() < need.value + 1024:
            self._ws = torch.empty(int(need.value) + 1024, dtype=torch.uint8, device=self.device)
        # the library wants a 1024-byte aligned workspace (TMA swizzle atoms); torch's allocator guarantees 512
        ptr = (self._ws.data_ptr() + 1023) & ~1023
        return ptr, self._ws.numel() - (ptr - self._ws.data_ptr())

    @staticmethod
    def _fmt(x):
        if x.dtype == torch.uint8:
            n, h, w, _ = x.shape
            return L.B2R_IN_U8_NHWC, n, h, w
        n, _, h, w = x.shape
        return L.B2R_IN_F32_NCHW, n, h, w

    @torch.no_grad()
    def restore(self, x: torch.Tensor, want_f32: bool = True, want_u8: bool = False):
        """b2r_unet_forward / b2r_resunet_forward on one resident batch; returns (f32 NCHW | None, u8 NHWC | None)."""
        fmt, n, h, w = self._fmt(x)
        x = x.contiguous()
        with torch.cuda.device(self.device):
            o32 = torch.empty((n, 3, h, w), dtype=torch.float32, device=self.device) if want_f32 else None
            o8 = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device) if want_u8 else None
            ws_ptr, ws_bytes = self._workspace(n, h, w)
            fn = L.load().b2r_unet_forward if self.arch == "simple_unet" else L.load().b2r_resunet_forward
            L.check(fn(self.handle, x.data_ptr(), fmt, None if o32 is None else o32.data_ptr(), None if o8 is None else o8.data_ptr(),
                       n, h, w, ws_ptr, ws_bytes, torch.cuda.current_stream().cuda_stream))
        return o32, o8

    @torch.no_grad()
    def classify(self, x: torch.Tensor, normalize: bool = True) -> torch.Tensor:
        """b2r_vgg16_forward: u8 NHWC (ToTensor + Normalize fused when normalize) or normalised f32 NCHW -> logits."""
        fmt, n, h, w = self._fmt(x)
        x = x.contiguous()
        with torch.cuda.device(self.device):
            logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=self.device)
            ws_ptr, ws_bytes = self._workspace(n, h, w)
            L.check(L.load().b2r_vgg16_forward(self.handle, x.data_ptr(), fmt, int(bool(normalize) and fmt == L.B2R_IN_U8_NHWC),
                                               logits.data_ptr(), n, h, w, ws_ptr, ws_bytes,
                                               torch.cuda.current_stream().cuda_stream))
        return logits
