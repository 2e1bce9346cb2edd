"""Generate the committed golden fixtures by RUNNING THE REFERENCE'S OWN CODE (imported from /root/reference).

    python tests/golden/make_golden.py          # needs /root/reference; writes tests/golden/*.npz

The reference cannot travel to the GPU box, so its outputs do: small inputs, the reference's outputs on them, and
the seeds that regenerate the synthetic checkpoints (the checkpoints themselves are too large to commit; a per-tensor
checksum is stored so that RNG drift would be detected).  Nothing here is imported by the product.
"""
from __future__ import annotations

import importlib.util
import io
import random
import sys
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def load_ref(filename: str):
    spec = importlib.util.spec_from_file_location("ref_" + filename.split("_")[0], REF / filename)
    mod = importlib.util.module_from_spec(spec)
    with redirect_stdout(io.StringIO()):      # 07_train_restoration.py prints at import (07:29-32)
        spec.loader.exec_module(mod)
    return mod


class ReplayNormal:
    """Stand-in for np.random.normal that returns a prepared array (the reference never seeds its RNG)."""

    def __init__(self, arr):
        self.arr = arr

    def __call__(self, loc, scale, size):
        assert tuple(size) == self.arr.shape
        return loc + scale * self.arr


class NpShim:
    """`np` as seen by a reference module: numpy with random.normal replaced (numpy itself is left untouched)."""

    def __init__(self, normal):
        import types
        self.random = types.SimpleNamespace(normal=normal)

    def __getattr__(self, name):
        return getattr(np, name)


class ReplayRandom:
    """Stand-in for the `random` module functions used by 14:38-55, replaying a fixed list of draws."""

    def __init__(self, draws):
        self.draws = list(draws)

    def _next(self):
        return self.draws.pop(0)

    def random(self):
        return self._next()

    def uniform(self, a, b):
        return self._next()

    def randint(self, a, b):
        return self._next()


def test_images(n=6, h=40, w=56, seed=0):
    rng = np.random.default_rng(seed)
    low = rng.integers(0, 256, (n, 6, 7, 3)).astype(np.float32)
    import cv2
    imgs = np.stack([cv2.resize(low[i], (w, h), interpolation=cv2.INTER_CUBIC) for i in range(n)])
    imgs = np.clip(imgs + rng.normal(0, 6, imgs.shape), 0, 255).astype(np.uint8)
    imgs[0, :4, :5] = 0        # saturated corners exercise the clip / REFLECT_101 paths
    imgs[1, -3:, -6:] = 255
    return imgs


def gen_degrade():
    import cv2
    r16 = load_ref("16_gen_compound_data.py")
    r14 = load_ref("14_train_unified_advanced.py")
    imgs = test_images()
    n, h, w, _ = imgs.shape
    rng = np.random.default_rng(1)
    z = rng.standard_normal((n, h, w, 3))                     # standard normals, scaled per call like np.random.normal
    # --- script 16
    out16 = []
    for i in range(n):
        r16.np = NpShim(ReplayNormal(z[i]))
        out16.append(r16.apply_compound_distortion(imgs[i]))
    # --- script 14 with explicit draw lists: [p_fog, intensity, u, p_noise, var, p_blur, degree, angle]
    cases = [
        dict(fog=(0.45, 1.1), noise=0.017, blur=(9, 133)),
        dict(fog=None, noise=0.03, blur=(5, 0)),
        dict(fog=(0.7, 1.2), noise=None, blur=(11, 271)),
        dict(fog=(0.3, 0.8), noise=0.01, blur=None),
        dict(fog=None, noise=None, blur=(15, 45)),
        dict(fog=None, noise=None, blur=None),
    ]
    out14, meta14 = [], []
    for i, c in enumerate(cases):
        draws = []
        if c["fog"]:
            draws += [0.1, c["fog"][0], c["fog"][1]]
        else:
            draws += [0.9]
        if c["noise"]:
            draws += [0.1, c["noise"]]
        else:
            draws += [0.9]
        if c["blur"]:
            draws += [0.1, c["blur"][0], c["blur"][1]]
        else:
            draws += [0.9]
        rr = ReplayRandom(draws)
        r14.random = rr
        r14.np = NpShim(ReplayNormal(z[i]))
        out14.append(r14.apply_random_distortions(imgs[i]))
        assert not rr.draws
        fog_t = (1.0 - c["fog"][0] * c["fog"][1]) if c["fog"] else np.nan
        meta14.append([fog_t, c["noise"] or np.nan, (c["blur"] or (0, 0))[0], (c["blur"] or (0, 0))[1]])
    # --- blur taps for every (degree, angle) the reference can draw, float32 as filter2D uses them
    taps = np.zeros((14, 361, 15, 15), np.float32)
    for d in range(2, 16):
        for a in range(361):
            M = cv2.getRotationMatrix2D((d / 2, d / 2), a, 1)
            k = cv2.warpAffine(np.diag(np.ones(d)), M, (d, d)) / d
            taps[d - 2, a, :d, :d] = k.astype(np.float32)
    np.savez_compressed(OUT / "degrade_ref.npz", images=imgs, z=z, out16=np.stack(out16), out14=np.stack(out14),
                        meta14=np.array(meta14, dtype=np.float64))
    np.savez_compressed(OUT / "blur_taps_ref.npz", taps=taps)
    # --- plain blur outputs (03's kernel, no normalise) on u8 for a spread of (d, angle): pins filter2D semantics
    blur_cases = [(2, 30), (5, 45), (7, 200), (10, 45), (11, 90), (12, 45), (15, 17)]
    blur_out = np.stack([cv2.filter2D(imgs[i % n], -1, cv2.warpAffine(
        np.diag(np.ones(d)), cv2.getRotationMatrix2D((d / 2, d / 2), a, 1), (d, d)) / d) for i, (d, a) in enumerate(blur_cases)])
    np.savez_compressed(OUT / "blur_ref.npz", images=imgs, cases=np.array(blur_cases), out=blur_out)


def checksum(sd):
    return {k: float(v.double().abs().sum()) for k, v in sd.items()}


def gen_models():
    from b200restore import synth
    r07 = load_ref("07_train_restoration.py")
    r17 = load_ref("17_run_unified_inference.py")
    import torchvision
    g = torch.Generator().manual_seed(5)
    res = {}
    with torch.no_grad():
        for arch, ctor, hw in (("simple_unet", r07.SimpleUNet, (16, 24)), ("resunet", r17.ResUNet, (16, 24))):
            sd = synth.synthetic_state_dict(arch, seed=11)
            m = ctor()
            m.load_state_dict(sd, strict=True)
            m.eval()
            x = torch.rand((2, 3) + hw, generator=g)
            res[arch + "_x"] = x.numpy()
            res[arch + "_y"] = m(x).numpy()
            cs = checksum(sd)
            res[arch + "_cs_keys"] = np.array(list(cs.keys()))
            res[arch + "_cs_vals"] = np.array(list(cs.values()))
        sd = synth.synthetic_state_dict("vgg16", seed=13)
        m = torchvision.models.vgg16(weights=None)
        m.classifier[6] = torch.nn.Linear(m.classifier[6].in_features, 43)   # 06_test_baseline.py:65-67
        m.load_state_dict(sd, strict=True)
        m.eval()
        x = torch.randn((2, 3, 64, 64), generator=g)
        res["vgg16_x"] = x.numpy()
        res["vgg16_y"] = m(x).numpy()
        cs = checksum(sd)
        res["vgg16_cs_keys"] = np.array(list(cs.keys()))
        res["vgg16_cs_vals"] = np.array(list(cs.values()))
    np.savez_compressed(OUT / "models_ref.npz", **res)


def load_ref_functions(filename, names):
    """Run ONLY the named top-level function definitions of a reference script that cannot be imported as a whole
    (13_pipeline_stress_test.py imports matplotlib at module level): the function bodies are the reference's own code,
    compiled from its own source file."""
    import ast
    import cv2
    import torch
    src = (REF / filename).read_text()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(keep) == len(names), [n.name for n in keep]
    ns = {"np": np, "cv2": cv2, "torch": torch}
    exec(compile(ast.Module(body=keep, type_ignores=[]), str(REF / filename), "exec"), ns)
    return ns


def gen_generators():
    """Single-degradation generators (02 / 03 / 04) and the stress-test distortions (13), random draws replayed."""
    r02 = load_ref("02_gen_noise.py")
    r03 = load_ref("03_gen_blur.py")
    r04 = load_ref("04_gen_fog.py")
    f13 = load_ref_functions("13_pipeline_stress_test.py", ["add_noise", "add_blur", "add_fog"])
    imgs = test_images()
    imgs[2] = np.clip(imgs[2].astype(np.int32) // 3 + 170, 0, 255).astype(np.uint8)   # bright image: script 02's "no negative value" branch
    imgs[3, 5:9, 5:9] = 0
    n, h, w, _ = imgs.shape
    rng = np.random.default_rng(7)
    z = rng.standard_normal((n, h, w, 3))
    # --- 02: var 0.02 as the script calls it, and a tiny variance on the bright image so that nothing goes negative
    var02 = np.array([0.02, 0.02, 0.0004, 0.02, 0.01, 0.05])
    out02 = []
    for i in range(n):
        r02.np = NpShim(ReplayNormal(z[i]))
        out02.append(r02.add_gaussian_noise(imgs[i], var=float(var02[i])))
    # --- 03: blur + joint min-max stretch
    cases03 = [(10, 45), (12, 45), (5, 0), (11, 200), (7, 90), (3, 135)]
    out03 = [r03.apply_motion_blur(imgs[i], degree=d, angle=a) for i, (d, a) in enumerate(cases03)]
    # --- 04: fog with the uniform draw replayed
    u04 = np.array([0.8, 1.2, 1.0, 0.93, 1.17, 0.85])
    inten04 = np.array([0.8, 0.8, 0.8, 0.5, 1.0, 0.1])
    out04 = []
    for i in range(n):
        r04.random = ReplayRandom([float(u04[i])])
        out04.append(r04.add_fog(imgs[i], fog_intensity=float(inten04[i])))
    # --- 13: Blur -> Fog -> Noise with u8 re-quantisation after every stage
    b13, f13o, n13 = [], [], []
    for i in range(n):
        f13["np"] = NpShim(ReplayNormal(z[i]))
        b = f13["add_blur"](imgs[i].copy())
        f = f13["add_fog"](b)
        b13.append(b)
        f13o.append(f)
        n13.append(f13["add_noise"](f))
    np.savez_compressed(OUT / "generators_ref.npz", images=imgs, z=z, var02=var02, out02=np.stack(out02),
                        cases03=np.array(cases03), out03=np.stack(out03), u04=u04, inten04=inten04,
                        out04=np.stack(out04), blur13=np.stack(b13), fog13=np.stack(f13o), noise13=np.stack(n13))


if __name__ == "__main__":
    assert REF.exists(), "the reference is not mounted; fixtures can only be generated where /root/reference exists"
    import sys as _sys
    if "--only-generators" not in _sys.argv:
        gen_degrade()
        gen_models()
    gen_generators()
    for p in sorted(OUT.glob("*.npz")):
        print(p.name, p.stat().st_size)
