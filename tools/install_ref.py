#!/usr/bin/env python
"""Install the UNMODIFIED reference scripts of the hot path into git-ignored `baseline/_ref/`.

    python tools/install_ref.py [--src /root/reference] [--check]

The reference is 19 flat Python scripts with no setup.py / pyproject, so `pip install --target baseline/_ref
/root/reference` has nothing to build ("neither 'setup.py' nor 'pyproject.toml' found"); the install IS a byte-for-byte
copy of the files the path needs.  `baseline/_ref/` is listed in .gitignore (the reference's sources never enter this
repo's history) but not in .gpurunignore, so it travels to the GPU box with the snapshot exactly like libb2r.so does;
`bench.py --impl reference` and `cpu_baseline` import the classes / functions from there with importlib (the file
names start with digits) and run them through their own code path.  MANIFEST.json records the sha256 of every file so
the bench can state that the arm it timed is the unmodified reference.

Files: 02/03/04 single degradations, 16 compound degradation, 14 random degradation + ResUNet,
07 SimpleUNet, 17 ResUNet + the batched inference loop, 18 the accuracy harness.  Scripts 08/13/15 import skimage /
matplotlib, which this image does not have; their hot-path code is a verbatim copy of 07/17's classes.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
DST = ROOT / "baseline" / "_ref"
FILES = ("02_gen_noise.py", "03_gen_blur.py", "04_gen_fog.py", "07_train_restoration.py",
         "14_train_unified_advanced.py", "16_gen_compound_data.py", "17_run_unified_inference.py",
         "18_test_unified_benchmark.py", "LICENSE")


def sha256(p: Path) -> str:
    return hashlib.sha256(p.read_bytes()).hexdigest()


def install(src: Path = Path("/root/reference"), dst: Path = DST) -> dict:
    if not src.is_dir():
        raise FileNotFoundError(f"reference not mounted at {src}")
    dst.mkdir(parents=True, exist_ok=True)
    manifest = {"source": str(src), "files": {}}
    for name in FILES:
        s = src / name
        if not s.exists():
            if name == "LICENSE":
                continue
            raise FileNotFoundError(s)
        shutil.copyfile(s, dst / name)
        manifest["files"][name] = sha256(dst / name)
    (dst / "MANIFEST.json").write_text(json.dumps(manifest, indent=1) + "\n")
    return manifest


def check(dst: Path = DST) -> bool:
    """True when every installed file still has the sha256 recorded at install time."""
    mf = dst / "MANIFEST.json"
    if not mf.exists():
        return False
    files = json.loads(mf.read_text())["files"]
    return all((dst / n).exists() and sha256(dst / n) == h for n, h in files.items())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    if a.check:
        ok = check()
        print("baseline/_ref: " + ("ok" if ok else "missing or modified"))
        sys.exit(0 if ok else 1)
    m = install(Path(a.src))
    print(f"installed {len(m['files'])} files into {DST}")


if __name__ == "__main__":
    main()
