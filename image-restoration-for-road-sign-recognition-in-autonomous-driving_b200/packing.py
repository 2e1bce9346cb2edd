"""Weight re-packing from the reference's state_dict layouts to the layouts the sm_100a kernels read.

Done once at `load_state_dict` time (or first forward): eval-mode BatchNorm folded into the preceding conv
(14_train_unified_advanced.py:100-104, eps = 1e-5), OIHW -> [C_out, K] bf16 with K ordered by k-block
(source, tap, 64-channel chunk), ConvTranspose2d [C_in, C_out, 2, 2] -> four stacked 1x1 matrices, VGG16
classifier[0] columns permuted from the NCHW flatten order (c*49 + h*7 + w) to the NHWC order the activations use.
Pure tensor reshuffling on whatever device the parameters live on; no arithmetic beyond the BN fold.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .ops import kblock

BN_EPS = 1e-5  # nn.BatchNorm2d default, used by ResidualBlock (14_train_unified_advanced.py:100)


def fold_bn(w: torch.Tensor, b: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor,
            mean: torch.Tensor, var: torch.Tensor, eps: float = BN_EPS) -> Tuple[torch.Tensor, torch.Tensor]:
    """BN(conv(x)) in eval mode == conv'(x) with w' = w * s[co], b' = (b - mean) * s + beta, s = gamma / sqrt(var+eps)."""
    w = w.double()
    s = gamma.double() / torch.sqrt(var.double() + eps)
    b0 = b.double() if b is not None else torch.zeros_like(s)
    w2 = w * s.view(-1, *([1] * (w.dim() - 1)))
    b2 = (b0 - mean.double()) * s + beta.double()
    return w2.float(), b2.float()


class KPlan:
    """Accumulates the K dimension of one fused layer: weight column groups + the matching k-block list.

    `finish()` gives the generic layout bf16 [C_out][K] (K = 64 x k-blocks).  For C_out = 64 layers `finish_w3()`
    additionally gives the tap-folded layout bf16 [192][64 x k-steps] read by csrc/conv_w3.cu: one k-step per
    (3x3 group, kernel row kh) whose rows are (kw, co), and one per centre block with the kw = 0, 2 rows zero."""

    def __init__(self, cout: int):
        self.cout = cout
        self.cols: List[torch.Tensor] = []
        self.kblocks: List[int] = []
        self._groups: List[Tuple[str, torch.Tensor]] = []   # ("3x3", w[co,64,3,3]) | ("1x1", w[co,64])

    def add_conv3x3(self, src: int, w: torch.Tensor, c_offset: int = 0) -> "KPlan":
        """w: [C_out, C_in_slice, 3, 3] applied to channels [c_offset, c_offset + C_in_slice) of source `src`."""
        co, ci, kh, kw = w.shape
        assert co == self.cout and kh == 3 and kw == 3 and ci % 64 == 0 and c_offset % 64 == 0
        # group order: 64-channel chunk outermost, then dw (kernel column), then dh (kernel row) — nine consecutive
        # k-blocks share one (source, chunk), which is the shape csrc/conv_n64.cu and csrc/conv_w3.cu recognise and
        # which keeps the generic kernel's nine shifted reads of a tile adjacent in L2
        for c in range(ci // 64):
            wc = w[:, c * 64:(c + 1) * 64]
            self._groups.append(("3x3", wc))
            for j in range(3):
                for i in range(3):
                    self.cols.append(wc[:, :, i, j])
                    self.kblocks.append(kblock(src, i - 1, j - 1, c_offset // 64 + c))
        return self

    def add_1x1(self, src: int, w: torch.Tensor, c_offset: int = 0) -> "KPlan":
        """w: [C_out, C_in_slice] (centre tap only) on channels [c_offset, ...) of source `src`."""
        co, ci = w.shape[:2]
        w = w.reshape(co, ci)
        assert co == self.cout and ci % 64 == 0 and c_offset % 64 == 0
        for c in range(ci // 64):
            wc = w[:, c * 64:(c + 1) * 64]
            self._groups.append(("1x1", wc))
            self.cols.append(wc)
            self.kblocks.append(kblock(src, 0, 0, c_offset // 64 + c))
        return self

    def finish(self, device=None) -> Tuple[torch.Tensor, List[int]]:
        wmat = torch.cat(self.cols, dim=1).to(torch.bfloat16).contiguous()
        if device is not None:
            wmat = wmat.to(device)
        return wmat, list(self.kblocks)

    def finish_w3(self, device=None) -> Optional[torch.Tensor]:
        if self.cout != 64:
            return None
        steps = []
        for kind, wc in self._groups:
            if kind == "3x3":
                for kh in range(3):                              # rows (kw, co), columns ci
                    steps.append(wc[:, :, kh, :].permute(2, 0, 1).reshape(192, 64))
            else:
                z = torch.zeros_like(wc)
                steps.append(torch.cat((z, wc, z), dim=0))       # only the kw = 1 (centre column) rows are non-zero
        w3 = torch.cat(steps, dim=1).to(torch.bfloat16).contiguous()
        return w3.to(device) if device is not None else w3


def pack_conv3x3(w: torch.Tensor, splits: Optional[Sequence[int]] = None):
    """Plain conv3x3 over one source, or over a channel concat of several sources (splits = channels per source)."""
    co, ci = w.shape[:2]
    splits = list(splits) if splits else [ci]
    assert sum(splits) == ci
    return plan_conv3x3(w, splits).finish()


def plan_conv3x3(w: torch.Tensor, splits: Optional[Sequence[int]] = None) -> KPlan:
    co, ci = w.shape[:2]
    splits = list(splits) if splits else [ci]
    assert sum(splits) == ci
    plan = KPlan(co)
    off = 0
    for s, n in enumerate(splits):
        plan.add_conv3x3(s, w[:, off:off + n])
        off += n
    return plan


def pack_convT2x2(w: torch.Tensor, b: torch.Tensor):
    """ConvTranspose2d(k=2, s=2) weight [C_in, C_out, 2, 2] -> [4*C_out, C_in] (quadrant q = 2*i + j major)."""
    ci, co = w.shape[:2]
    wm = w.permute(2, 3, 1, 0).reshape(4 * co, ci)  # [(i, j, co), ci]
    bias = b.repeat(4)
    return wm.to(torch.bfloat16).contiguous(), bias.float().contiguous()


def pack_fc_from_nchw_flatten(w: torch.Tensor, c: int, h: int, wd: int) -> torch.Tensor:
    """Linear weight whose columns index an NCHW flatten -> columns for the NHWC flatten of the same tensor."""
    o = w.shape[0]
    return w.reshape(o, c, h, wd).permute(0, 2, 3, 1).reshape(o, h * wd * c).to(torch.bfloat16).contiguous()


def pack_conv_c3(w: torch.Tensor) -> torch.Tensor:
    """First-layer weights f32 [64, 3, 3, 3] (OIHW) -> bf16 [64, 64] for csrc/conv_c3.cu: column 2*t + j (j = 0, 1) holds
    w[co][ci][kh][kw] with t = (kh*3 + kw)*3 + ci (each weight twice: the kernel feeds activations as hi/lo bf16 pairs)."""
    assert tuple(w.shape) == (64, 3, 3, 3)
    flat = w.permute(0, 2, 3, 1).reshape(64, 27)            # [co][(kh, kw, ci)]
    out = torch.zeros((64, 64), dtype=torch.float32, device=w.device)
    out[:, 0:54:2] = flat
    out[:, 1:54:2] = flat
    return out.to(torch.bfloat16).contiguous()
