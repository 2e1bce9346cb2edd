"""Guard bands around every output of the conv / element-wise kernels (compute-sanitizer is not available on the GPU pool, so
out-of-bounds WRITES are caught here): each output tensor is a window inside a larger buffer pre-filled with 0xA5 bytes; after
the launch the bytes before and after the window must be untouched and the window must be fully written (no 0xA5A5 left in a
bf16 output: that pattern is -7.2e-17 as bf16 and never the result of these layers).  Shapes are picked so that tiles are
clipped at the right / bottom edge and the last tile pair is incomplete (odd tile counts)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

PAD = 4096


class Guarded:
    def __init__(self, shape, dtype):
        self.nbytes = int(torch.tensor([], dtype=dtype).element_size()) * int(torch.Size(shape).numel())
        self.buf = torch.full((PAD + self.nbytes + PAD,), 0xA5, dtype=torch.uint8, device="cuda")
        self.t = self.buf[PAD:PAD + self.nbytes].view(dtype).view(shape)

    def check(self, what, written=True):
        torch.cuda.synchronize()
        assert bool((self.buf[:PAD] == 0xA5).all()), f"{what}: bytes BEFORE the output were overwritten"
        assert bool((self.buf[PAD + self.nbytes:] == 0xA5).all()), f"{what}: bytes AFTER the output were overwritten"
        if written and self.t.dtype == torch.bfloat16:
            left = int((self.t.view(torch.int16) == torch.tensor(0xA5A5 - 65536, dtype=torch.int16)).sum())
            assert left == 0, f"{what}: {left} output elements were never written"


def _plan(packing, co, srcs_c, shortcut=None, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    ci = sum(srcs_c)
    w = torch.randn((co, ci, 3, 3), generator=g) * (2.0 / (9 * ci)) ** 0.5
    plan = packing.KPlan(co)
    off = 0
    for s, c in enumerate(srcs_c):
        plan.add_conv3x3(s, w[:, off:off + c])
        off += c
    if shortcut is not None:
        plan.add_1x1(len(srcs_c), torch.randn((co, shortcut), generator=g) * (1.0 / shortcut) ** 0.5)
    wm, kbl = plan.finish()
    return wm.cuda(), kbl, (plan.finish_w3().cuda() if co == 64 else None)


def _src(n, h, w, c, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn((n, h, w, c), generator=g) * 0.5).to(torch.bfloat16).cuda()


@pytest.mark.parametrize("n,h,w,srcs_c,co,pooled,flags", [
    (3, 40, 40, (64,), 64, False, 0),        # tap-folded pair mode: 3 x 5 = 15 tiles per image, 45 tiles (odd), clipped columns / rows
    (3, 40, 40, (64,), 64, True, 0),         # + fused pool
    (1, 24, 30, (64, 64), 64, False, 0),     # two sources (cat), W not a multiple of 14
    (3, 40, 40, (64,), 64, False, 8),        # B2R_CONV_NO_PAIR: single-CTA tap-folded kernel
    (3, 24, 40, (64,), 128, False, 0),       # pair + halo N = 128, clipped tiles
    (3, 24, 40, (64,), 128, True, 0),
    (1, 24, 40, (128,), 128, False, 8),      # single-CTA halo kernel
    (3, 12, 12, (128,), 256, False, 0),      # pair N = 256, tiles spanning images, odd tile count
    (3, 12, 12, (128,), 256, True, 0),
    (3, 12, 12, (128,), 256, False, 8),      # single-CTA N = 256 (two epilogue groups)
    (5, 6, 6, (64,), 128, False, 4),         # B2R_CONV_NO_HALO: generic N = 128 (two co-resident CTAs, two epilogue groups)
    (3, 12, 12, (128,), 256, False, 12),     # NO_PAIR | NO_HALO: generic N = 256, two epilogue groups, one staging buffer each
    (3, 12, 12, (128,), 256, True, 12),
])
def test_conv_outputs_stay_inside_their_buffers(n, h, w, srcs_c, co, pooled, flags):
    from b200restore import ops, packing, _lib as L
    wm, kbl, w3 = _plan(packing, co, srcs_c)
    srcs = [_src(n, h, w, c, 10 + i) for i, c in enumerate(srcs_c)]
    out = Guarded((n, h, w, co), torch.bfloat16)
    pool = Guarded((n, h // 2, w // 2, co), torch.bfloat16) if pooled else None
    ops.conv_gemm(srcs, wm, torch.zeros(co, device="cuda"), kbl, act=L.B2R_ACT_RELU, out=out.t,
                  out_pool=pool.t if pooled else None, weights_w3=w3, flags=flags)
    what = f"conv {srcs_c}->{co} @{h}x{w} n={n} flags={flags} ({(L.load().b2r_last_conv_kernel() or b'?').decode()})"
    print(what)
    out.check(what)
    if pooled:
        pool.check(what + " pooled")


@pytest.mark.parametrize("n,h,w,ci,co", [(3, 10, 12, 64, 64), (1, 7, 9, 128, 64), (2, 5, 6, 256, 128)])
def test_conv_transpose_outputs_stay_inside_their_buffers(n, h, w, ci, co):
    from b200restore import ops, packing, _lib as L
    g = torch.Generator(device="cpu").manual_seed(3)
    wm, bias = packing.pack_convT2x2(torch.randn((ci, co, 2, 2), generator=g) * (1.0 / ci) ** 0.5, torch.zeros(co))
    out = Guarded((n, 2 * h, 2 * w, co), torch.bfloat16)
    ops.conv_gemm([_src(n, h, w, ci, 4)], wm.cuda(), bias.cuda(), None, out=out.t, out_mode=L.B2R_OUT_CONVT2X2)
    out.check(f"convT {ci}->{co} @{h}x{w} ({(L.load().b2r_last_conv_kernel() or b'?').decode()})")


def test_fused_head_outputs_stay_inside_their_buffers():
    from b200restore import ops, packing, _lib as L
    n, h, w = 3, 40, 40
    wm, kbl, w3 = _plan(packing, 64, (64,), shortcut=128)
    srcs = [_src(n, h, w, 64, 1), _src(n, h, w, 128, 2)]
    g = torch.Generator(device="cpu").manual_seed(5)
    o32, o8 = Guarded((n, 3, h, w), torch.float32), Guarded((n, h, w, 3), torch.uint8)
    ops.conv_gemm(srcs, wm, torch.zeros(64, device="cuda"), kbl, act=L.B2R_ACT_RELU, weights_w3=w3,
                  head_w=(torch.randn((3, 64), generator=g) * 0.1).cuda(), head_b=torch.zeros(3, device="cuda"),
                  head_out_f32=o32.t, head_out_u8=o8.t)
    o32.check("head f32")
    o8.check("head u8")


@pytest.mark.parametrize("fmt", ["u8", "f32"])
def test_first_layer_and_small_kernels_stay_inside_their_buffers(fmt):
    from b200restore import ops, packing, _lib as L
    n, h, w = 3, 20, 36                                        # clipped 8 x 16 tiles in both directions
    g = torch.Generator(device="cpu").manual_seed(6)
    wp, bias = packing.pack_conv_c3(torch.randn((64, 3, 3, 3), generator=g) * 0.2), torch.zeros(64)
    x = (torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8) if fmt == "u8"
         else torch.rand((n, 3, h, w), generator=g)).cuda()
    out = Guarded((n, h, w, 64), torch.bfloat16)
    ops.conv3x3_c3(x, wp.cuda(), bias.cuda(), act=L.B2R_ACT_RELU, normalize=fmt == "u8", out=out.t)
    out.check(f"conv3x3_c3 {fmt}")
    src = _src(n, 21, 37, 64, 7)                               # odd map: floor semantics
    pooled = Guarded((n, 10, 18, 64), torch.bfloat16)
    ops.maxpool2x2(src, out=pooled.t)
    pooled.check("maxpool2x2 odd map")
    rs = Guarded((n, 23, 35, 64), torch.bfloat16)
    ops.resize_nearest(src, 23, 35, out=rs.t)
    rs.check("resize_nearest")
