#!/usr/bin/env python
"""Compile the UNMODIFIED reference extension (fairnr/clib/src/*.cpp, *.cu) into oracle/_ref/ref_ext.so.

TEST INFRASTRUCTURE ONLY.  Nothing here is product code: the sources are compiled where they
lie under /root/reference (never copied into this repo) and only the build OUTPUT lands in the
git-ignored directory oracle/_ref/, which still travels to the GPU box with `gpurun`.

The recipe is a direct nvcc/g++ invocation (we do not run the reference's setup.py):
  * -O2 only, as the reference's setup.py:25-28 does, plus the sm_100a gencode it lacks;
  * `assert` stays active (no -DNDEBUG), like a stock reference build;
  * TORCH_EXTENSION_NAME=ref_ext so `import ref_ext` exposes the 7 functions of
    fairnr/clib/src/binding.cpp:11-20.

On the GPU box /root/reference does not exist; this script then only reports whether a
prebuilt oracle/_ref/ref_ext.so is present.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("NSVF_REFERENCE", "/root/reference")
CLIB = os.path.join(REF, "fairnr", "clib")
SO = os.path.join(OUT, "ref_ext.so")


def build(force=False, verbose=True):
    if not os.path.isdir(CLIB):
        if verbose:
            print("[oracle/_ref] %s absent; using prebuilt %s: %s" % (CLIB, SO, os.path.exists(SO)))
        return os.path.exists(SO)
    srcs = sorted(os.path.join(CLIB, "src", f) for f in os.listdir(os.path.join(CLIB, "src"))
                  if f.endswith((".cpp", ".cu")))
    if os.path.exists(SO) and not force and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in srcs):
        return True
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    inc = ["-I" + os.path.join(CLIB, "include")]
    for p in ce.include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]:
        inc += ["-isystem", p]
    defs = ["-DTORCH_EXTENSION_NAME=ref_ext", "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    objs, cmds = [], []
    for s in srcs:
        o = os.path.join(OUT, os.path.basename(s) + ".o")
        objs.append(o)
        if s.endswith(".cu"):
            cmds.append(["nvcc", "-c", s, "-o", o, "-O2", "-std=c++17", "--expt-relaxed-constexpr",
                         "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                         "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                         "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                         "-w"] + defs + inc)
        else:
            cmds.append(["g++", "-c", s, "-o", o, "-O2", "-std=c++17", "-fPIC", "-w"] + defs + inc)

    def run(c):
        r = subprocess.run(c, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference build failed: %s\n%s" % (" ".join(c), r.stderr[-4000:]))
    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, cmds))
    libdir = ce.library_paths()[0]
    link = ["g++", "-shared", "-o", SO] + objs + ["-L" + libdir, "-L/usr/local/cuda/lib64",
            "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
            "-Wl,-rpath," + libdir]
    run(link)
    for o in objs:
        os.remove(o)
    if verbose:
        print("[oracle/_ref] built", SO)
    return True


def load():
    """Import oracle/_ref/ref_ext.so as a module (needs torch imported first). Returns None if absent."""
    if not os.path.exists(SO):
        return None
    import importlib.util
    import torch  # noqa: F401  (ref_ext links against libtorch)
    spec = importlib.util.spec_from_file_location("ref_ext", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    sys.exit(0 if ok else 1)
