"""GPU parity of the tcgen05 implicit-GEMM conv (b2r_conv_gemm) against PyTorch fp32 convs on the same bf16-rounded
operands.  Floating point: the comparator is torch fp32 (TF32 off); tolerance = bf16 output rounding (2^-8 relative)
plus fp32 accumulation-order noise, stated per test.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _ops():
    from b200restore import ops, packing, _lib
    return ops, packing, _lib


def nhwc_bf16(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def to_nchw_f32(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def assert_close_bf16(got, ref, what, rel=2.0 ** -7, abs_=2e-2):
    """|got - ref| <= abs_ + rel*|ref| elementwise; bf16 rounding of the output alone is 2^-9 relative."""
    err = (got - ref).abs()
    bound = abs_ + rel * ref.abs()
    bad = err > bound
    assert not bad.any(), (f"{what}: {int(bad.sum())} / {bad.numel()} outside tolerance; max err "
                           f"{float(err.max()):.4g} at ref {float(ref.flatten()[err.argmax()]):.4g}")


@pytest.mark.parametrize("n,h,w,ci,co,block_n,tile", [
    (2, 16, 32, 64, 64, 0, (0, 0, 0)),       # smallest: K = 9 blocks, one N tile
    (1, 24, 40, 128, 128, 0, (16, 8, 1)),    # partial tiles at the right / bottom edge
    (3, 8, 8, 64, 256, 256, (8, 8, 2)),      # images packed into one M tile, odd batch (clipped in N)
    (8, 4, 4, 128, 64, 0, (4, 4, 8)),        # tiny maps: 8 images per tile
    (2, 14, 14, 256, 512, 256, (0, 0, 0)),   # VGG conv5-like 14x14, K = 36 blocks, two N tiles
    (1, 16, 16, 512, 128, 128, (16, 8, 1)),  # K = 72 blocks
    (1, 32, 32, 64, 128, 64, (0, 0, 0)),     # block_n smaller than C_out -> several N tiles
])
def test_conv3x3_relu(n, h, w, ci, co, block_n, tile):
    ops, packing, L = _ops()
    x = rnd(n, ci, h, w, seed=1)
    wt = rnd(co, ci, 3, 3, scale=(2.0 / (9 * ci)) ** 0.5, seed=2)
    b = rnd(co, scale=0.1, seed=3)
    xb = nhwc_bf16(x)
    wm, kbl = packing.pack_conv3x3(wt)
    out = torch.full((n, h, w, co), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.conv_gemm([xb], wm.cuda(), b, kbl, act=L.B2R_ACT_RELU, out=out, block_n=block_n, tile=tile)
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(to_nchw_f32(xb), wt.to(torch.bfloat16).float(), b, padding=1))
    assert_close_bf16(to_nchw_f32(out), ref, "conv3x3+relu")


def test_conv3x3_two_sources_prelu_pool():
    """Concat-free decoder conv: cat((a, s), 1) -> conv3x3 -> PReLU, with the fused 2x2 max-pool second output."""
    ops, packing, L = _ops()
    n, h, w = 2, 16, 32
    a = rnd(n, 64, h, w, seed=4)
    s = rnd(n, 128, h, w, seed=5)
    wt = rnd(128, 192, 3, 3, scale=(2.0 / (9 * 192)) ** 0.5, seed=6)
    b = rnd(128, scale=0.1, seed=7)
    ab, sb = nhwc_bf16(a), nhwc_bf16(s)
    wm, kbl = packing.pack_conv3x3(wt, splits=[64, 128])
    out = torch.full((n, h, w, 128), float("nan"), dtype=torch.bfloat16, device="cuda")
    pool = torch.full((n, h // 2, w // 2, 128), float("nan"), dtype=torch.bfloat16, device="cuda")
    slope = 0.25
    ops.conv_gemm([ab, sb], wm.cuda(), b, kbl, act=L.B2R_ACT_PRELU, slope=slope, out=out, out_pool=pool)
    torch.cuda.synchronize()
    xin = torch.cat((to_nchw_f32(ab), to_nchw_f32(sb)), 1)
    ref = F.prelu(F.conv2d(xin, wt.to(torch.bfloat16).float(), b, padding=1), torch.tensor([slope], device="cuda"))
    assert_close_bf16(to_nchw_f32(out), ref, "two-source conv")
    # the pooled output must be exactly the max-pool of the stored full-resolution output
    ref_pool = F.max_pool2d(to_nchw_f32(out), 2, 2)
    assert torch.equal(to_nchw_f32(pool), ref_pool)


def test_pool_only_output():
    ops, packing, L = _ops()
    n, h, w = 1, 16, 16
    x = rnd(n, 64, h, w, seed=8)
    wt = rnd(64, 64, 3, 3, scale=(2.0 / 576) ** 0.5, seed=9)
    b = rnd(64, scale=0.1, seed=10)
    xb = nhwc_bf16(x)
    wm, kbl = packing.pack_conv3x3(wt)
    pool = torch.full((n, h // 2, w // 2, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.conv_gemm([xb], wm.cuda(), b, kbl, act=L.B2R_ACT_RELU, out_pool=pool)
    torch.cuda.synchronize()
    ref = F.max_pool2d(F.relu(F.conv2d(to_nchw_f32(xb), wt.to(torch.bfloat16).float(), b, padding=1)), 2, 2)
    assert_close_bf16(to_nchw_f32(pool), ref, "pool-only conv")


def test_residual_block_second_conv():
    """relu(conv3x3(y) + conv1x1(x) + bias): the ResidualBlock tail as one K loop over two sources."""
    ops, packing, L = _ops()
    n, h, w = 2, 8, 16
    y = rnd(n, 128, h, w, seed=11)
    x = rnd(n, 64, h, w, seed=12)
    w2 = rnd(128, 128, 3, 3, scale=(1.0 / (9 * 128)) ** 0.5, seed=13)
    ws = rnd(128, 64, 1, 1, scale=(1.0 / 64) ** 0.5, seed=14)
    b = rnd(128, scale=0.1, seed=15)
    yb, xb = nhwc_bf16(y), nhwc_bf16(x)
    plan = packing.KPlan(128).add_conv3x3(0, w2).add_1x1(1, ws)
    wm, kbl = plan.finish()
    out = torch.empty((n, h, w, 128), dtype=torch.bfloat16, device="cuda")
    ops.conv_gemm([yb, xb], wm.cuda(), b, kbl, act=L.B2R_ACT_RELU, out=out)
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(to_nchw_f32(yb), w2.to(torch.bfloat16).float(), b, padding=1)
                 + F.conv2d(to_nchw_f32(xb), ws.to(torch.bfloat16).float()))
    assert_close_bf16(to_nchw_f32(out), ref, "residual tail")


def test_identity_shortcut_is_exact_add():
    """An identity 1x1 k-block adds the bf16 block input exactly (fp32 accumulate)."""
    ops, packing, L = _ops()
    n, h, w = 1, 8, 16
    x = rnd(n, 64, h, w, seed=16)
    xb = nhwc_bf16(x)
    zero3 = torch.zeros(64, 64, 3, 3, device="cuda")
    plan = packing.KPlan(64).add_conv3x3(0, zero3).add_1x1(0, torch.eye(64, device="cuda"))
    wm, kbl = plan.finish()
    out = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device="cuda")
    ops.conv_gemm([xb], wm.cuda(), torch.zeros(64, device="cuda"), kbl, act=L.B2R_ACT_NONE, out=out)
    torch.cuda.synchronize()
    assert torch.equal(out, xb)


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 8, 16, 256, 128), (1, 4, 4, 64, 64), (3, 14, 14, 128, 64)])
def test_conv_transpose_2x2(n, h, w, ci, co):
    ops, packing, L = _ops()
    x = rnd(n, ci, h, w, seed=17)
    wt = rnd(ci, co, 2, 2, scale=(1.0 / ci) ** 0.5, seed=18)
    b = rnd(co, scale=0.1, seed=19)
    xb = nhwc_bf16(x)
    wm, bias4 = packing.pack_convT2x2(wt, b)
    out = torch.full((n, 2 * h, 2 * w, co), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.conv_gemm([xb], wm.cuda(), bias4.cuda(), None, act=L.B2R_ACT_NONE, out=out, out_mode=L.B2R_OUT_CONVT2X2)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(to_nchw_f32(xb), wt.to(torch.bfloat16).float(), b, stride=2)
    assert_close_bf16(to_nchw_f32(out), ref, "convT 2x2")


@pytest.mark.parametrize("bsz,k,o", [(200, 512, 256), (64, 25088, 128), (1, 4096, 64)])
def test_linear_as_conv(bsz, k, o):
    ops, packing, L = _ops()
    x = rnd(bsz, k, seed=20)
    wt = rnd(o, k, scale=(1.0 / k) ** 0.5, seed=21)
    b = rnd(o, scale=0.1, seed=22)
    xb = x.to(torch.bfloat16).contiguous()
    out = torch.full((1, 1, bsz, o), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.conv_gemm([xb.view(1, 1, bsz, k)], wt.to(torch.bfloat16).contiguous(), b, None, act=L.B2R_ACT_RELU, out=out)
    torch.cuda.synchronize()
    ref = F.relu(xb.float() @ wt.to(torch.bfloat16).float().t() + b)
    assert_close_bf16(out.view(bsz, o).float(), ref, "linear")


def test_many_tiles_persistent_loop():
    """More tiles than SMs, several waves, both accumulator stages and all smem stages wrap many times."""
    ops, packing, L = _ops()
    n, h, w, ci, co = 16, 64, 64, 64, 64
    x = rnd(n, ci, h, w, seed=23)
    wt = rnd(co, ci, 3, 3, scale=(2.0 / 576) ** 0.5, seed=24)
    b = rnd(co, scale=0.1, seed=25)
    xb = nhwc_bf16(x)
    wm, kbl = packing.pack_conv3x3(wt)
    out = torch.empty((n, h, w, co), dtype=torch.bfloat16, device="cuda")
    pool = torch.empty((n, h // 2, w // 2, co), dtype=torch.bfloat16, device="cuda")
    ops.conv_gemm([xb], wm.cuda(), b, kbl, act=L.B2R_ACT_RELU, out=out, out_pool=pool)
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(to_nchw_f32(xb), wt.to(torch.bfloat16).float(), b, padding=1))
    assert_close_bf16(to_nchw_f32(out), ref, "many tiles")
    assert torch.equal(to_nchw_f32(pool), F.max_pool2d(to_nchw_f32(out), 2, 2))


def test_argument_errors():
    ops, packing, L = _ops()
    xb = torch.zeros((1, 8, 16, 64), dtype=torch.bfloat16, device="cuda")
    wm = torch.zeros((64, 576), dtype=torch.bfloat16, device="cuda")
    b = torch.zeros(64, device="cuda")
    _, kbl = packing.pack_conv3x3(torch.zeros(64, 64, 3, 3))
    with pytest.raises(L.B2RError):
        ops.conv_gemm([xb.cpu()], wm, b, kbl, out=torch.empty_like(xb))           # CPU tensor: no fallback
    with pytest.raises(L.B2RError):
        ops.conv_gemm([xb], wm, b, kbl)                                            # no output
    with pytest.raises(L.B2RError):
        ops.conv_gemm([xb], wm, b, kbl, out=torch.empty_like(xb), tile=(16, 4, 1))  # tile != 128 pixels
    with pytest.raises(L.B2RError):
        odd = torch.zeros((1, 7, 16, 64), dtype=torch.bfloat16, device="cuda")
        ops.conv_gemm([odd], wm, b, kbl, out_pool=torch.empty((1, 3, 8, 64), dtype=torch.bfloat16, device="cuda"))


# ------------------------------------------------------------------------------------------------------------------
# C_out = 64 specialisation (csrc/conv_n64.cu): resident weights + column-shifted halo boxes
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,h,w,splits,shortcut,pool,tile", [
    (2, 32, 32, (64,), None, False, (0, 0, 0)),          # plain 64 -> 64
    (1, 224, 224, (64,), None, True, (0, 0, 0)),         # VGG conv1_2 shape with the fused pool
    (3, 40, 24, (64,), None, False, (0, 0, 0)),          # partial tiles on both edges
    (2, 48, 64, (64, 64), None, False, (16, 8, 1)),      # dec1.c1: cat of two sources, 16 x 8 tile
    (2, 32, 48, (64,), "identity", True, (0, 0, 0)),     # res1.c2: conv + identity shortcut + pool
    (2, 32, 32, (64,), (64, 128), False, (8, 16, 1)),    # dec2.c2: conv(y) + 1x1 shortcut over cat(64, 128)
    (2, 40, 56, (64, 128), None, False, (0, 0, 0)),      # dec2.c1: 192 -> 64, weights (216 KB) streamed through a ring
    (1, 112, 112, (64, 128), None, True, (0, 0, 0)),     # same at the real size, with the fused pool
])
def test_conv_n64_specialisation_equals_generic_kernel(n, h, w, splits, shortcut, pool, tile):
    """Same k-block order => same accumulation order: the specialised kernel must reproduce the generic kernel's
    bf16 output bit for bit, and both must match the fp32 torch reference within bf16 tolerance."""
    ops, packing, L = _ops()
    srcs = [nhwc_bf16(rnd(n, c, h, w, seed=40 + i)) for i, c in enumerate(splits)]
    ci = sum(splits)
    wt = rnd(64, ci, 3, 3, scale=(2.0 / (9 * ci)) ** 0.5, seed=50)
    b = rnd(64, scale=0.1, seed=51)
    plan = packing.KPlan(64)
    off = 0
    for s, c in enumerate(splits):
        plan.add_conv3x3(s, wt[:, off:off + c])
        off += c
    ref = F.conv2d(torch.cat([to_nchw_f32(s) for s in srcs], 1), wt.to(torch.bfloat16).float(), b, padding=1)
    all_srcs = list(srcs)
    if shortcut == "identity":
        plan.add_1x1(len(all_srcs), torch.eye(64, device="cuda"))
        xs = nhwc_bf16(rnd(n, 64, h, w, seed=52))
        all_srcs.append(xs)
        ref = ref + to_nchw_f32(xs)
    elif shortcut is not None:
        ws = rnd(64, sum(shortcut), 1, 1, scale=(1.0 / sum(shortcut)) ** 0.5, seed=53)
        off = 0
        sc_in = []
        for c in shortcut:
            xs = nhwc_bf16(rnd(n, c, h, w, seed=54 + c))
            plan.add_1x1(len(all_srcs), ws[:, off:off + c])
            all_srcs.append(xs)
            sc_in.append(to_nchw_f32(xs))
            off += c
        ref = ref + F.conv2d(torch.cat(sc_in, 1), ws.to(torch.bfloat16).float())
    ref = F.relu(ref)
    wm, kbl = plan.finish()
    wm, w3 = wm.cuda(), plan.finish_w3().cuda()
    outs = []
    for flags in (L.B2R_CONV_NO_W3, L.B2R_CONV_GENERIC_ONLY, 0):
        out = torch.full((n, h, w, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
        pl = torch.full((n, h // 2, w // 2, 64), float("nan"), dtype=torch.bfloat16, device="cuda") if pool else None
        ops.conv_gemm(all_srcs, wm, b, kbl, act=L.B2R_ACT_RELU, out=out, out_pool=pl, weights_w3=w3,
                      tile=tile if flags == L.B2R_CONV_NO_W3 else (0, 0, 0), flags=flags)
        torch.cuda.synchronize()
        outs.append((out, pl))
    assert_close_bf16(to_nchw_f32(outs[0][0]), ref, "n64 kernel vs torch")
    assert torch.equal(outs[0][0], outs[1][0]), "resident-weight (n64) and generic kernels disagree"
    # tap-folded kernel (conv_w3.cu): different fp32 summation order (three kw partial sums added in the epilogue),
    # so equal to the others only up to one bf16 rounding of the output
    assert_close_bf16(to_nchw_f32(outs[2][0]), ref, "w3 kernel vs torch")
    d = (outs[2][0].float() - outs[0][0].float()).abs()
    assert bool((d <= 2.0 ** -7 * outs[0][0].float().abs() + 1e-3).all()), float(d.max())
    if pool:
        assert torch.equal(outs[0][1], outs[1][1])
        for o, pl in outs:
            assert torch.equal(to_nchw_f32(pl), F.max_pool2d(to_nchw_f32(o), 2, 2))


def test_conv_w3_full_size_and_odd_widths():
    """224 x 224 (16 x 28 tiles of 14 x 8, no waste) and widths that are not multiples of 14 (clipped tiles)."""
    ops, packing, L = _ops()
    for n, h, w in ((2, 224, 224), (1, 40, 30), (3, 16, 100), (1, 8, 14)):
        x = nhwc_bf16(rnd(n, 64, h, w, seed=70))
        wt = rnd(64, 64, 3, 3, scale=(2.0 / 576) ** 0.5, seed=71)
        b = rnd(64, scale=0.1, seed=72)
        plan = packing.plan_conv3x3(wt)
        wm, kbl = plan.finish()
        out = torch.full((n, h, w, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
        pl = torch.full((n, h // 2, w // 2, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
        ops.conv_gemm([x], wm.cuda(), b, kbl, act=L.B2R_ACT_PRELU, slope=0.3, out=out, out_pool=pl,
                      weights_w3=plan.finish_w3().cuda())
        torch.cuda.synchronize()
        ref = F.prelu(F.conv2d(to_nchw_f32(x), wt.to(torch.bfloat16).float(), b, padding=1),
                      torch.tensor([0.3], device="cuda"))
        assert_close_bf16(to_nchw_f32(out), ref, f"w3 {n}x{h}x{w}")
        assert torch.equal(to_nchw_f32(pl), F.max_pool2d(to_nchw_f32(out), 2, 2))


def test_conv_n64_many_tiles_and_ring_wraps():
    ops, packing, L = _ops()
    n, h, w = 24, 112, 112
    x = nhwc_bf16(rnd(n, 64, h, w, seed=60))
    wt = rnd(64, 64, 3, 3, scale=(2.0 / 576) ** 0.5, seed=61)
    b = rnd(64, scale=0.1, seed=62)
    wm, kbl = packing.pack_conv3x3(wt)
    out = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device="cuda")
    ops.conv_gemm([x], wm.cuda(), b, kbl, act=L.B2R_ACT_RELU, out=out)
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(to_nchw_f32(x), wt.to(torch.bfloat16).float(), b, padding=1))
    assert_close_bf16(to_nchw_f32(out), ref, "n64 many tiles")


def test_conv_w3_random_shapes_against_generic_kernel():
    """Seeded sweep over ragged shapes / source splits / pooling for the tap-folded kernel (two MMA issuer warps, per-warp
    staging, store warp): every output must agree with the generic kernel's within one bf16 rounding, pooled outputs
    must be the exact 2x2 maxima of the stored tile, and nothing outside the tensors may be touched (NaN guards)."""
    ops, packing, L = _ops()
    g = torch.Generator().manual_seed(1234)
    for case in range(10):
        n = int(torch.randint(1, 5, (1,), generator=g))
        h = 2 * int(torch.randint(1, 40, (1,), generator=g))
        w = 2 * int(torch.randint(1, 60, (1,), generator=g))
        splits = [(64,), (64, 64), (128,), (64, 128)][int(torch.randint(0, 4, (1,), generator=g))]
        pool = bool(torch.randint(0, 2, (1,), generator=g))
        store_full = bool(torch.randint(0, 2, (1,), generator=g)) or not pool
        srcs = [nhwc_bf16(rnd(n, c, h, w, seed=300 + 7 * case + i)) for i, c in enumerate(splits)]
        ci = sum(splits)
        wt = rnd(64, ci, 3, 3, scale=(2.0 / (9 * ci)) ** 0.5, seed=400 + case)
        b = rnd(64, scale=0.1, seed=500 + case)
        plan = packing.KPlan(64)
        off = 0
        for s, c in enumerate(splits):
            plan.add_conv3x3(s, wt[:, off:off + c])
            off += c
        wm, kbl = plan.finish()
        wm, w3 = wm.cuda(), plan.finish_w3().cuda()
        res = []
        for flags in (L.B2R_CONV_GENERIC_ONLY, 0):
            # guard rows around the outputs catch out-of-bounds stores
            out_g = torch.full((n + 2, h, w, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
            pl_g = torch.full((n + 2, h // 2, w // 2, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
            out, pl = out_g[1:n + 1], pl_g[1:n + 1]
            ops.conv_gemm(srcs, wm, b, kbl, act=L.B2R_ACT_PRELU, slope=0.25, out=out if (store_full or flags) else None,
                          out_pool=pl if pool else None, weights_w3=w3, flags=flags)
            torch.cuda.synchronize()
            assert bool(torch.isnan(out_g[0]).all()) and bool(torch.isnan(out_g[n + 1]).all()), f"case {case}: OOB store"
            assert bool(torch.isnan(pl_g[0]).all()) and bool(torch.isnan(pl_g[n + 1]).all()), f"case {case}: OOB pool store"
            res.append((out.clone(), pl.clone()))
        ref_out, ref_pl = res[0]
        out, pl = res[1]
        if store_full:
            d = (out.float() - ref_out.float()).abs()
            assert bool((d <= 2.0 ** -7 * ref_out.float().abs() + 1e-3).all()), (case, n, h, w, splits, float(d.max()))
        if pool:
            exact = F.max_pool2d(to_nchw_f32(out), 2, 2) if store_full else None
            if exact is not None:
                assert torch.equal(to_nchw_f32(pl), exact), (case, n, h, w, splits)
            d = (pl.float() - ref_pl.float()).abs()
            assert bool((d <= 2.0 ** -7 * ref_pl.float().abs() + 1e-3).all()), (case, "pool", float(d.max()))


@pytest.mark.parametrize("n,h,w,splits,co,shortcut,pool", [
    (2, 32, 48, (64,), 128, None, False),          # VGG conv2_1 shape class (N = 128, two co-resident CTAs)
    (1, 112, 112, (128,), 128, None, True),        # conv2_2 with the fused pool
    (2, 56, 56, (128, 256), 128, None, False),     # dec3.c1: concat of two sources
    (2, 40, 24, (128,), 256, None, False),         # N = 256, ragged tiles
    (3, 28, 28, (256,), 256, None, True),          # 32 x 4 tiles
    (2, 14, 14, (512,), 512, None, False),         # two n-tiles per pixel tile
    (2, 32, 32, (128,), 128, (64,), True),         # res2.c2: conv + 1x1 shortcut (centre k-blocks) + pool
])
def test_halo_mode_equals_per_tap_mode(n, h, w, splits, co, shortcut, pool):
    """Halo mode (one (TH + 2)-row box per (chunk, dw), A and B in separate rings) issues the same MMAs in the same order
    as the per-tap mode: outputs must be bit-identical, and both within bf16 tolerance of the fp32 torch reference."""
    ops, packing, L = _ops()
    srcs = [nhwc_bf16(rnd(n, c, h, w, seed=600 + i)) for i, c in enumerate(splits)]
    ci = sum(splits)
    wt = rnd(co, ci, 3, 3, scale=(2.0 / (9 * ci)) ** 0.5, seed=610)
    b = rnd(co, scale=0.1, seed=611)
    plan = packing.KPlan(co)
    off = 0
    for s, c in enumerate(splits):
        plan.add_conv3x3(s, wt[:, off:off + c])
        off += c
    ref = F.conv2d(torch.cat([to_nchw_f32(s) for s in srcs], 1), wt.to(torch.bfloat16).float(), b, padding=1)
    all_srcs = list(srcs)
    if shortcut is not None:
        ws = rnd(co, sum(shortcut), 1, 1, scale=(1.0 / sum(shortcut)) ** 0.5, seed=612)
        xs = nhwc_bf16(rnd(n, shortcut[0], h, w, seed=613))
        plan.add_1x1(len(all_srcs), ws)
        all_srcs.append(xs)
        ref = ref + F.conv2d(to_nchw_f32(xs), ws.to(torch.bfloat16).float())
    ref = F.relu(ref)
    wm, kbl = plan.finish()
    wm = wm.cuda()
    outs = []
    for flags in (0, L.B2R_CONV_NO_HALO):
        out = torch.full((n + 2, h, w, co), float("nan"), dtype=torch.bfloat16, device="cuda")
        pl = torch.full((n + 2, h // 2, w // 2, co), float("nan"), dtype=torch.bfloat16, device="cuda") if pool else None
        ops.conv_gemm(all_srcs, wm, b, kbl, act=L.B2R_ACT_RELU, out=out[1:n + 1], out_pool=None if pl is None else pl[1:n + 1],
                      flags=flags)
        torch.cuda.synchronize()
        assert bool(torch.isnan(out[0]).all()) and bool(torch.isnan(out[n + 1]).all()), "store outside the tensor"
        outs.append((out[1:n + 1].clone(), None if pl is None else pl[1:n + 1].clone()))
    assert_close_bf16(to_nchw_f32(outs[0][0]), ref, "halo mode vs torch")
    assert torch.equal(outs[0][0], outs[1][0]), "halo and per-tap modes disagree"
    if pool:
        assert torch.equal(outs[0][1], outs[1][1])
        assert torch.equal(to_nchw_f32(outs[0][1]), F.max_pool2d(to_nchw_f32(outs[0][0]), 2, 2))


@pytest.mark.parametrize("n,h,w,ci,co,pool", [
    (2, 16, 16, 64, 256, False),        # one pair per n-tile, tiny
    (3, 56, 56, 128, 256, False),       # conv3_1 shape class, 8x8x2 tiles, odd number of pixel tiles
    (2, 28, 28, 256, 512, True),        # two n-tiles, fused pool
    (5, 14, 14, 512, 512, False),       # tiles spanning several images, odd tile count
    (1, 40, 24, 64, 256, False),        # ragged tiles
])
def test_pair_mode_equals_single_cta_mode(n, h, w, ci, co, pool):
    """cta_group::2 pair mode (M = 256 across two CTAs, each loading half of the weights) must reproduce the single-CTA
    kernel bit for bit (same k-block order, fp32 accumulation per output element in the same order)."""
    ops, packing, L = _ops()
    x = nhwc_bf16(rnd(n, ci, h, w, seed=700))
    wt = rnd(co, ci, 3, 3, scale=(2.0 / (9 * ci)) ** 0.5, seed=701)
    b = rnd(co, scale=0.1, seed=702)
    wm, kbl = packing.plan_conv3x3(wt).finish()
    wm = wm.cuda()
    ref = F.relu(F.conv2d(to_nchw_f32(x), wt.to(torch.bfloat16).float(), b, padding=1))
    outs = []
    for flags in (0, L.B2R_CONV_NO_PAIR):
        out = torch.full((n + 2, h, w, co), float("nan"), dtype=torch.bfloat16, device="cuda")
        pl = torch.full((n + 2, h // 2, w // 2, co), float("nan"), dtype=torch.bfloat16, device="cuda") if pool else None
        ops.conv_gemm([x], wm, b, kbl, act=L.B2R_ACT_RELU, out=out[1:n + 1], out_pool=None if pl is None else pl[1:n + 1],
                      flags=flags)
        torch.cuda.synchronize()
        assert bool(torch.isnan(out[0]).all()) and bool(torch.isnan(out[n + 1]).all()), "store outside the tensor"
        outs.append((out[1:n + 1].clone(), None if pl is None else pl[1:n + 1].clone()))
    assert_close_bf16(to_nchw_f32(outs[0][0]), ref, "pair mode vs torch")
    assert torch.equal(outs[0][0], outs[1][0]), "pair and single-CTA modes disagree"
    if pool:
        assert torch.equal(outs[0][1], outs[1][1])


def test_pair_and_halo_modes_random_shapes():
    """Seeded sweep over batch / size / channel splits / pooling / shortcuts for the C_out = 128 (halo mode) and
    C_out = 256, 512 (cta_group::2 pair mode) kernels: bit-identical to the per-tap single-CTA kernel, NaN guards intact."""
    ops, packing, L = _ops()
    g = torch.Generator().manual_seed(4321)
    ri = lambda lo, hi: int(torch.randint(lo, hi, (1,), generator=g))   # noqa: E731
    for case in range(12):
        co = (128, 256, 512)[ri(0, 3)]
        n, h, w = ri(1, 6), 2 * ri(2, 30), 2 * ri(2, 30)
        splits = [(64,), (128,), (64, 64), (128, 64), (256,)][ri(0, 5)]
        pool = bool(ri(0, 2))
        shortcut = bool(ri(0, 2))
        srcs = [nhwc_bf16(rnd(n, c, h, w, seed=800 + 11 * case + i)) for i, c in enumerate(splits)]
        ci = sum(splits)
        wt = rnd(co, ci, 3, 3, scale=(2.0 / (9 * ci)) ** 0.5, seed=900 + case)
        b = rnd(co, scale=0.1, seed=950 + case)
        plan = packing.KPlan(co)
        off = 0
        for s, c in enumerate(splits):
            plan.add_conv3x3(s, wt[:, off:off + c])
            off += c
        all_srcs = list(srcs)
        if shortcut:
            xs = nhwc_bf16(rnd(n, 64, h, w, seed=990 + case))
            plan.add_1x1(len(all_srcs), rnd(co, 64, 1, 1, scale=0.1, seed=995 + case))
            all_srcs.append(xs)
        wm, kbl = plan.finish()
        wm = wm.cuda()
        res = []
        for flags in (0, L.B2R_CONV_NO_HALO | L.B2R_CONV_NO_PAIR):
            out = torch.full((n + 2, h, w, co), float("nan"), dtype=torch.bfloat16, device="cuda")
            pl = torch.full((n + 2, h // 2, w // 2, co), float("nan"), dtype=torch.bfloat16, device="cuda")
            ops.conv_gemm(all_srcs, wm, b, kbl, act=L.B2R_ACT_PRELU, slope=0.2, out=out[1:n + 1],
                          out_pool=pl[1:n + 1] if pool else None, flags=flags)
            torch.cuda.synchronize()
            assert bool(torch.isnan(out[0]).all()) and bool(torch.isnan(out[n + 1]).all()), (case, "OOB store")
            assert bool(torch.isnan(pl[0]).all()) and bool(torch.isnan(pl[n + 1]).all()), (case, "OOB pool store")
            assert not bool(torch.isnan(out[1:n + 1]).any()), (case, "unwritten output")
            res.append((out[1:n + 1].clone(), pl[1:n + 1].clone()))
        assert torch.equal(res[0][0], res[1][0]), (case, co, n, h, w, splits, pool, shortcut)
        if pool:
            assert torch.equal(res[0][1], res[1][1]), (case, "pool")


@pytest.mark.parametrize("n,h,w,splits,pool,shortcut", [
    (1, 16, 8, (64,), False, False),           # a single tile pair
    (3, 24, 40, (64,), False, False),          # odd number of pixel tiles: the last pair's second CTA stores nothing
    (2, 112, 112, (128,), True, False),        # VGG conv2_2 with the fused pool
    (5, 56, 56, (128, 256), False, True),      # dec3 second-conv shape class: two sources + centre k-blocks, odd image count
    (1, 8, 8, (64,), True, False),             # one clipped tile only: falls back to a single-CTA kernel
])
def test_pair_halo_mode_equals_halo_mode(n, h, w, splits, pool, shortcut):
    """Round 2: C_out = 128 layers run as cta_group::2 pairs ON halo boxes (conv_gemm_pairhalo_kernel<128>).  Same MMAs in the
    same order as the single-CTA halo kernel (B2R_CONV_NO_PAIR): bit-identical outputs, nothing stored outside the tensor."""
    ops, packing, L = _ops()
    co = 128
    srcs = [nhwc_bf16(rnd(n, c, h, w, seed=1200 + i)) for i, c in enumerate(splits)]
    ci = sum(splits)
    wt = rnd(co, ci, 3, 3, scale=(2.0 / (9 * ci)) ** 0.5, seed=1210)
    b = rnd(co, scale=0.1, seed=1211)
    plan = packing.KPlan(co)
    off = 0
    for s, c in enumerate(splits):
        plan.add_conv3x3(s, wt[:, off:off + c])
        off += c
    all_srcs = list(srcs)
    if shortcut:
        plan.add_1x1(len(all_srcs), rnd(co, 64, 1, 1, scale=0.1, seed=1212))
        all_srcs.append(nhwc_bf16(rnd(n, 64, h, w, seed=1213)))
    wm, kbl = plan.finish()
    wm = wm.cuda()
    res, kernels = [], []
    for flags in (0, L.B2R_CONV_NO_PAIR, L.B2R_CONV_NO_PAIR | L.B2R_CONV_NO_HALO):
        out = torch.full((n + 2, h, w, co), float("nan"), dtype=torch.bfloat16, device="cuda")
        pl = torch.full((n + 2, h // 2, w // 2, co), float("nan"), dtype=torch.bfloat16, device="cuda")
        ops.conv_gemm(all_srcs, wm, b, kbl, act=L.B2R_ACT_RELU, out=out[1:n + 1], out_pool=pl[1:n + 1] if pool else None,
                      flags=flags)
        kernels.append(L.load().b2r_last_conv_kernel().decode())
        torch.cuda.synchronize()
        assert bool(torch.isnan(out[0]).all()) and bool(torch.isnan(out[n + 1]).all()), "store outside the tensor"
        assert bool(torch.isnan(pl[0]).all()) and bool(torch.isnan(pl[n + 1]).all()), "pool store outside the tensor"
        assert not bool(torch.isnan(out[1:n + 1]).any()), "unwritten output"
        res.append((out[1:n + 1].clone(), pl[1:n + 1].clone()))
    if n * h * w > 128:
        assert kernels[0] == "conv_gemm_pairhalo_kernel<128>" and kernels[1] == "conv_gemm_halo_kernel<128>", kernels
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][0], res[2][0]), kernels
    if pool:
        assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][1], res[2][1])


def test_conv_w3_pair_mode_equals_single_cta_mode():
    """Round 2: the tap-folded kernel as cta_group::2 pairs (two neighbouring tiles per MMA of M = 256 x N = 192, each CTA
    holding half of the weight rows) must reproduce the single-CTA kernel (B2R_CONV_NO_PAIR) bit for bit: same k-steps in the
    same order.  Covers odd tile counts (the last pair's second CTA walks one tile past the end), pooled and head outputs,
    two sources, 1x1 centre groups (N = 64 MMAs on 32 + 32 weight rows), a first group that is a centre group, and the
    192 -> 64 layer whose weights only stay resident in pair mode (the single-CTA kernel streams them)."""
    ops, packing, L = _ops()
    g = torch.Generator().manual_seed(77)
    ri = lambda lo, hi: int(torch.randint(lo, hi, (1,), generator=g))   # noqa: E731
    cases = [(1, 8, 14, (64,), (), False, False), (1, 8, 28, (64,), (), False, False), (3, 24, 42, (64,), (), True, False),
             (2, 224, 224, (64,), (64,), True, False), (2, 112, 112, (64, 128), (), False, False),
             (2, 56, 70, (64,), (64, 128), False, False), (3, 40, 30, (64,), (64, 64), False, True),
             (1, 16, 16, (), (64,), False, False)]
    for _ in range(6):
        cases.append((ri(1, 5), 2 * ri(1, 40), 2 * ri(1, 60), [(64,), (64, 64), (128,)][ri(0, 3)], [(), (64,), (128,)][ri(0, 3)],
                      bool(ri(0, 2)), False))
    for ci_, (n, h, w, splits, sc, pool, head) in enumerate(cases):
        srcs = [nhwc_bf16(rnd(n, c, h, w, seed=1400 + 13 * ci_ + i)) for i, c in enumerate(splits)]
        plan = packing.KPlan(64)
        if splits:
            ci = sum(splits)
            wt = rnd(64, ci, 3, 3, scale=(2.0 / (9 * ci)) ** 0.5, seed=1500 + ci_)
            off = 0
            for s, c in enumerate(splits):
                plan.add_conv3x3(s, wt[:, off:off + c])
                off += c
        for c in sc:   # 1x1 centre groups over extra sources (ResidualBlock shortcut); with no 3x3 group it comes first
            srcs.append(nhwc_bf16(rnd(n, c, h, w, seed=1600 + 17 * ci_ + len(srcs))))
            plan.add_1x1(len(srcs) - 1, rnd(64, c, 1, 1, scale=(1.0 / c) ** 0.5, seed=1700 + ci_ + len(srcs)))
        b = rnd(64, scale=0.1, seed=1800 + ci_)
        wm, kbl = plan.finish()
        wm, w3 = wm.cuda(), plan.finish_w3().cuda()
        hw_ = rnd(3, 64, scale=0.1, seed=1900 + ci_).contiguous() if head else None
        hb_ = rnd(3, scale=0.1, seed=1950 + ci_) if head else None
        res, kernels = [], []
        for flags in (0, L.B2R_CONV_NO_PAIR):
            out_g = torch.full((n + 2, h, w, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
            pl_g = torch.full((n + 2, h // 2, w // 2, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
            h32 = torch.full((n + 2, 3, h, w), float("nan"), device="cuda") if head else None
            h8 = torch.full((n + 2, h, w, 3), 77, dtype=torch.uint8, device="cuda") if head else None
            ops.conv_gemm(srcs, wm, b, kbl, act=L.B2R_ACT_PRELU, slope=0.25, out=None if head else out_g[1:n + 1],
                          out_pool=pl_g[1:n + 1] if pool else None, weights_w3=w3, flags=flags, head_w=hw_, head_b=hb_,
                          head_out_f32=None if h32 is None else h32[1:n + 1], head_out_u8=None if h8 is None else h8[1:n + 1])
            kernels.append(L.load().b2r_last_conv_kernel().decode())
            torch.cuda.synchronize()
            assert bool(torch.isnan(out_g[0]).all()) and bool(torch.isnan(out_g[n + 1]).all()), (ci_, "OOB store")
            assert bool(torch.isnan(pl_g[0]).all()) and bool(torch.isnan(pl_g[n + 1]).all()), (ci_, "OOB pool store")
            if head:
                assert bool(torch.isnan(h32[0]).all()) and bool(torch.isnan(h32[n + 1]).all()), (ci_, "OOB head store")
                assert bool((h8[0] == 77).all()) and bool((h8[n + 1] == 77).all())
                assert not bool(torch.isnan(h32[1:n + 1]).any())
                res.append((h32[1:n + 1].clone(), h8[1:n + 1].clone()))
            else:
                assert not bool(torch.isnan(out_g[1:n + 1]).any()), (ci_, "unwritten output")
                res.append((out_g[1:n + 1].clone(), pl_g[1:n + 1].clone()))
        tiles = n * ((h + 7) // 8) * ((w + 13) // 14)
        if tiles >= 2:
            assert "pair" in kernels[0] and "pair" not in kernels[1], kernels
        assert torch.equal(res[0][0], res[1][0]), (ci_, n, h, w, splits, sc, kernels)
        if pool or head:
            assert torch.equal(res[0][1], res[1][1]), (ci_, "second output")


@pytest.mark.parametrize("co,with_w3", [(64, True), (64, False), (128, False), (256, False)])
@pytest.mark.parametrize("act,slope", [("none", 0.0), ("relu", 0.0), ("prelu", 0.25), ("prelu", 1.0), ("prelu", -0.3), ("prelu", 1.7)])
def test_every_activation_form(co, with_w3, act, slope):
    """The epilogues pick one of four forms per launch (conv_common.cuh::act_form): plain pack, ReLU on the packed pair,
    max(x, slope x) for 0 <= slope <= 1, and the general max(x,0) + slope min(x,0) for any other slope (a trained PReLU may
    leave [0, 1]).  All of them against torch on the tap-folded kernel (co = 64 with the folded weights) and the generic family."""
    ops, packing, L = _ops()
    n, h, w, ci = 2, 16, 28, 64
    x = rnd(n, ci, h, w, seed=21)
    wt = rnd(co, ci, 3, 3, scale=(2.0 / (9 * ci)) ** 0.5, seed=22)
    b = rnd(co, scale=0.3, seed=23)
    plan = packing.plan_conv3x3(wt.cpu())
    wm, kbl = plan.finish()
    w3 = plan.finish_w3().cuda() if with_w3 else None
    out = torch.empty((n, h, w, co), dtype=torch.bfloat16, device="cuda")
    code = {"none": L.B2R_ACT_NONE, "relu": L.B2R_ACT_RELU, "prelu": L.B2R_ACT_PRELU}[act]
    ops.conv_gemm([nhwc_bf16(x)], wm.cuda(), b, kbl, act=code, slope=slope, out=out, weights_w3=w3)
    ref = F.conv2d(nhwc_bf16(x).float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), b, padding=1)
    if act == "relu":
        ref = F.relu(ref)
    elif act == "prelu":
        ref = F.prelu(ref, torch.tensor([slope], device="cuda"))
    assert_close_bf16(to_nchw_f32(out), ref, f"{act} slope {slope} co {co} ({(L.load().b2r_last_conv_kernel() or b'?').decode()})")
