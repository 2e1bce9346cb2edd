"""Build recipe for libb2r.so (in-tree, sm_100a only).

`python -m b200restore.build` or `__graft_entry__.build()` runs it.  nvcc cross-compiles without a GPU; the
shared object has no link-time dependency on libcuda (driver entry points are resolved at run time), so it can be
dlopen'ed on a CPU-only box for the symbol check in tests/test_abi.py.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libb2r.so"
SOURCES = ["api_common.cu", "conv_gemm.cu", "conv_n64.cu", "conv_w3.cu", "conv_c3.cu", "elementwise.cu", "degrade.cu", "generators.cu", "ssim.cu", "net_plan.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=default",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (looked at $NVCC, PATH, /usr/local/cuda/bin/nvcc)")


def have_nvcc() -> bool:
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    deps.append(PKG_DIR.parent / "include" / "b2r.h")
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: Path | None = None) -> Path:
    """Compile every CUDA source into libb2r.so; returns the path.

    `defines` / `out` build a VARIANT next to the product library (e.g. defines=("B2R_TIMELINE",) for the role
    timelines of tools/role_timeline.py, loaded through $B2R_LIB); the product build takes neither."""
    variant = bool(defines) or out is not None
    lib_path = Path(out) if out is not None else LIB_PATH
    if variant and out is None:
        raise ValueError("a variant build needs its own output path")
    if not variant and not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    objdir = PKG_DIR / "build" / ("variant_" + "_".join(defines) if variant else "")
    objdir.mkdir(parents=True, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = objdir / (src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib_path), *objs,
            "-cudart", "static"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return lib_path


if __name__ == "__main__":
    # python -m b200restore.build [--force] [-v] [--define NAME ... --out PATH]
    argv = sys.argv[1:]
    defs = tuple(argv[i + 1] for i, a in enumerate(argv) if a == "--define")
    outp = next((Path(argv[i + 1]) for i, a in enumerate(argv) if a == "--out"), None)
    path = build(force="--force" in argv, verbose="-v" in argv, defines=defs, out=outp)
    print(path)
