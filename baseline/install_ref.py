#!/usr/bin/env python
"""Install the UNMODIFIED reference into the git-ignored directory baseline/_ref/ so that it travels to the GPU box.

BASELINE INFRASTRUCTURE ONLY — nothing under baseline/ is imported by the product (nsvf_b200/).  Two steps, both
reading /root/reference where it lies (read-only) and writing only under baseline/_ref/ (listed in .gitignore, not
in .gpurunignore):

  1. the base contract's offline install, which runs the reference's OWN setup.py (CUDAExtension 'fairnr.clib._ext',
     -O2) and therefore yields the reference's clib kernels compiled for sm_100a:
         pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse
                     --target baseline/_ref  <copy of /root/reference under /tmp>
     (a /tmp copy because the build writes into the source tree; --no-deps because fairseq & co. are not
     installable offline).  The reference's setup() lists no `packages=` — it is meant for `pip install --editable` —
     so this installs ONLY fairnr/clib/_ext.*.so;
  2. what the editable install would have exposed: the reference's Python package fairnr/ (data, modules, models,
     clib/__init__.py, LICENSE: MIT), copied verbatim next to the extension.  Never edited; loaded through
     baseline/ref_loader.py, which stubs the packages that are not installed (fairseq, plyfile, ...).

On the GPU box /root/reference does not exist: the script then only reports whether baseline/_ref is populated.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("NSVF_REFERENCE", "/root/reference")


def ext_path():
    so = glob.glob(os.path.join(OUT, "fairnr", "clib", "_ext*.so"))
    return so[0] if so else None


def installed():
    return ext_path() is not None and os.path.exists(os.path.join(OUT, "fairnr", "modules", "renderer.py"))


def install(force=False, verbose=True):
    if not os.path.isdir(os.path.join(REF, "fairnr")):
        if verbose:
            print("[baseline/_ref] %s absent; prebuilt install present: %s" % (REF, installed()))
        return installed()
    if installed() and not force:
        return True
    os.makedirs(OUT, exist_ok=True)
    if ext_path() is None or force:
        tmp = "/tmp/nsvf_ref_copy"
        shutil.rmtree(tmp, ignore_errors=True)
        shutil.copytree(REF, tmp, ignore=shutil.ignore_patterns(".git"))
        env = dict(os.environ, TORCH_CUDA_ARCH_LIST="10.0a", MAX_JOBS="8")
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--upgrade", "--target", OUT, tmp]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference install failed:\n" + r.stdout[-2000:] + r.stderr[-4000:])
        shutil.rmtree(tmp, ignore_errors=True)
    # step 2: the Python package, verbatim (py files + LICENSE)
    for root, dirs, files in os.walk(os.path.join(REF, "fairnr")):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", "src", "include")]
        rel = os.path.relpath(root, REF)
        os.makedirs(os.path.join(OUT, rel), exist_ok=True)
        for f in files:
            if f.endswith(".py"):
                shutil.copy2(os.path.join(root, f), os.path.join(OUT, rel, f))
    shutil.copy2(os.path.join(REF, "LICENSE"), os.path.join(OUT, "LICENSE"))
    if verbose:
        print("[baseline/_ref] installed:", ext_path())
    return installed()


if __name__ == "__main__":
    sys.exit(0 if install(force="--force" in sys.argv) else 1)
