"""Host logic of nsvf_b200.blas (the cuBLAS selection for the field MLP's contractions); no GPU needed: the toolkit's
libcublas loads without one.  Each case runs in a fresh interpreter because the selection has to precede `import torch`."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, env=None):
    e = dict(os.environ)
    e.pop("CUBLAS_EMULATE_SINGLE_PRECISION", None)
    e.update(env or {})
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env=e, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_selection_precedes_torch_and_binds_the_toolkit_library():
    r = _run("import json, os, sys\n"
             "from nsvf_b200 import blas\n"
             "assert 'torch' not in sys.modules, 'importing nsvf_b200.blas must not import torch'\n"
             "ok = blas.use_system_cublas()\n"
             "again = blas.use_system_cublas()\n"
             "import torch\n"
             "maps = open('/proc/self/maps').read()\n"
             "print(json.dumps({'ok': ok, 'again': again, 'mode': blas.mode(), 'env': os.environ.get('CUBLAS_EMULATE_SINGLE_PRECISION'),\n"
             "                  'toolkit_mapped': '/usr/local/cuda' in maps and 'libcublas.so' in maps}))")
    if not r["ok"]:      # machine without cuBLAS >= 12.9: the module must say so and leave the environment alone
        assert "SGEMM" in r["mode"] and r["env"] is None
        return
    assert r["again"] is True and r["env"] == "1" and r["toolkit_mapped"] and "BF16x9" in r["mode"]


def test_selection_after_torch_import_is_an_error_and_opt_out_works():
    r = _run("import json, torch\n"
             "from nsvf_b200 import blas\n"
             "try:\n"
             "    blas.use_system_cublas(); err = ''\n"
             "except RuntimeError as e:\n"
             "    err = str(e)\n"
             "print(json.dumps({'err': err, 'emulated': blas.emulated()}))")
    assert "before `import torch`" in r["err"] and r["emulated"] is False
    r = _run("import json, os\n"
             "from nsvf_b200 import blas\n"
             "print(json.dumps({'ok': blas.use_system_cublas(), 'env': os.environ.get('CUBLAS_EMULATE_SINGLE_PRECISION')}))",
             env={"NSVF_NO_CUBLAS_EMULATION": "1"})
    assert r["ok"] is False and r["env"] is None
