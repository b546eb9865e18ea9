"""TEST / BASELINE INFRASTRUCTURE ONLY - never imported by the product (pyvoxeldosimetry_b200).

Build and import the REAL reference (devhliu/PyVoxelDosimetry, /root/reference) for (a) pinning the oracle
(oracle/gen_golden.py), (b) the reference arm of bench.py (`--impl reference` drives the reference's own
KernelConvolutionCalculator.calculate_dose_rate, core/kernel_convolution.py:48-76) and (c) running the
reference's own example scripts unmodified against the drop-in (tests/test_ref_examples.py).

Recipe; nothing from the reference enters the repository history:
  build_ref()       copies /root/reference/pyvoxeldosimetry and /root/reference/examples/*.py byte for byte into
                    oracle/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box like a built .so) and
                    writes oracle/_ref/MANIFEST.json with the sha256 of every copied file;
  import_reference() registers permissive stub modules for the third-party imports this image lacks (nibabel,
                    SimpleITK, pydicom, cupy, matplotlib - the kernel-convolution path touches none of them), puts
                    oracle/_ref first on sys.path and imports the reference package under its own name;
                    the Lu177/F18/... kernel JSON files contain `//` comments (Lu177/Lu177.json:16) - stripped
                    before json.loads; save_kernel's png/npy side effects are switched off.
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_SRC = "/root/reference"
REF_DIR = os.path.join(HERE, "_ref")
STUBBED = ["nibabel", "nibabel.processing", "SimpleITK", "pydicom", "pydicom.dataset", "pydicom.uid", "cupy", "matplotlib",
           "matplotlib.pyplot"]


class Stub(types.ModuleType):
    """Permissive stand-in for an absent third-party module: attributes are child stubs, calls return the stub."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = Stub(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:  # decorators such as @cp.fuse() hand their argument back
            return a[0]
        return self

    def __iter__(self):  # `fig, ax = plt.subplots()` style unpacking in example scripts
        return iter((self, self))


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "MANIFEST.json"))


def build_ref(force: bool = False) -> str:
    """Copy the reference package and its example scripts into oracle/_ref (needs /root/reference: this container only)."""
    if not os.path.isdir(os.path.join(REF_SRC, "pyvoxeldosimetry")):
        raise FileNotFoundError(f"{REF_SRC} is not present (GPU box): oracle/_ref must be built in the build container")
    if available() and not force:
        return REF_DIR
    shutil.rmtree(REF_DIR, ignore_errors=True)
    os.makedirs(REF_DIR)
    shutil.copytree(os.path.join(REF_SRC, "pyvoxeldosimetry"), os.path.join(REF_DIR, "pyvoxeldosimetry"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.makedirs(os.path.join(REF_DIR, "examples"))
    for f in sorted(os.listdir(os.path.join(REF_SRC, "examples"))):
        if f.endswith(".py"):
            shutil.copyfile(os.path.join(REF_SRC, "examples", f), os.path.join(REF_DIR, "examples", f))
    manifest = {}
    for root, _, files in os.walk(REF_DIR):
        for f in sorted(files):
            p = os.path.join(root, f)
            manifest[os.path.relpath(p, REF_DIR)] = _sha(p)
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_SRC, "files": manifest}, fh, indent=0, sort_keys=True)
    return REF_DIR


def verify_against_source() -> list:
    """-> files of oracle/_ref that differ from /root/reference (empty = byte identical); needs both trees."""
    with open(os.path.join(REF_DIR, "MANIFEST.json")) as fh:
        files = json.load(fh)["files"]
    bad = []
    for rel, sha in files.items():
        src = os.path.join(REF_SRC, rel)
        if not os.path.exists(src) or _sha(src) != sha or _sha(os.path.join(REF_DIR, rel)) != sha:
            bad.append(rel)
    return bad


def install_stubs() -> None:
    for name in STUBBED:
        if name not in sys.modules:
            sys.modules[name] = Stub(name)


def import_reference(root: str | None = None):
    """Import the real reference package (as `pyvoxeldosimetry`) from `root` (default oracle/_ref).  The repo's own
    alias package of the same name is taken off sys.path / sys.modules first.  Returns the package module."""
    root = root or REF_DIR
    if not os.path.isdir(os.path.join(root, "pyvoxeldosimetry")):
        raise FileNotFoundError(f"no reference package under {root}: run oracle/ref_loader.py (build_ref) in the build container")
    install_stubs()
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
    for k in [k for k in sys.modules if k == "pyvoxeldosimetry" or k.startswith("pyvoxeldosimetry.")]:
        del sys.modules[k]
    sys.path.insert(0, root)
    import pyvoxeldosimetry  # noqa: F401  (the reference)
    from pyvoxeldosimetry.data.dose_kernels import base_kernel

    def _load_config(self, config_path):
        with open(config_path, "r") as f:
            return json.loads(re.sub(r"//[^\n]*", "", f.read()))

    base_kernel.BaseKernelGenerator._load_config = _load_config
    base_kernel.BaseKernelGenerator.save_kernel = lambda self, kernel, output_dir: None  # no png / npy side effects
    sys.path.append(REPO)  # `oracle` and the repo's other top-level modules stay importable (after the reference)
    return pyvoxeldosimetry


if __name__ == "__main__":
    d = build_ref(force="--force" in sys.argv)
    bad = verify_against_source()
    print(f"{d}: {len(json.load(open(os.path.join(d, 'MANIFEST.json')))['files'])} files, {len(bad)} differ from {REF_SRC}")
    sys.exit(1 if bad else 0)
