"""Build recipe: compile the UNMODIFIED reference (daviyan5/ninpol, Cython/OpenMP) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path; the product
(`ninpol_b200`) never imports it.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may use the result.

What this does
  * copies the reference's `ninpol/` package sources from /root/reference to a scratch dir in /tmp
    (the reference tree is read-only and its sources must not enter this repository),
  * cythonizes the six extensions the reference's setup.py:22-61 lists with the release compiler
    directives of setup.py:100-108 (boundscheck/wraparound/nonecheck/initializedcheck off,
    cdivision=True) and the release flags of setup.py:89-91 (-O3 -fopenmp),
  * installs ONLY build outputs into oracle/_ref/ninpol/ (the .so files, plus the few tiny
    runtime files the package cannot import without: the package's pure-Python stubs
    (__init__.py, utils/common.py, ...) and utils/point_ordering.yaml).  oracle/_ref/ is git-ignored, but it does travel to the GPU box.

`import ninpol` additionally needs `meshio` (interpolator.pyx:8), which is not installed in this
image; oracle/shim/meshio.py is a minimal stand-in (Mesh/CellBlock containers only).

Usage:  python oracle/build_ref.py [--force]
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("NINPOL_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

_SETUP = r'''
import os, numpy as np
from setuptools import setup, Extension
from Cython.Build import cythonize
names = [("_interpolator", "interpolator", None), ("_interpolator", "grid", "c++"),
         ("_methods", "idw", None), ("_methods", "gls", "c++"), ("_methods", "ls", None),
         ("_interpolator", "logger", None)]
exts = []
for sub, mod, lang in names:
    kw = dict(language=lang) if lang else {}
    exts.append(Extension(name=f"ninpol.{sub}.{mod}",
                          sources=[os.path.join("ninpol", sub, mod + ".pyx")],
                          extra_compile_args=["-O3", "-fopenmp"], extra_link_args=["-fopenmp"],
                          define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")],
                          include_dirs=[np.get_include(), os.path.join("ninpol", "utils")], **kw))
directives = dict(boundscheck=False, wraparound=False, nonecheck=False, initializedcheck=False,
                  cdivision=True, profile=False, linetrace=False)
setup(name="ninpol", packages=[], ext_modules=cythonize(exts, language_level="3",
      nthreads=os.cpu_count(), compiler_directives=directives))
'''


def is_built():
    need = ["_interpolator/interpolator", "_interpolator/grid", "_interpolator/logger",
            "_methods/idw", "_methods/ls", "_methods/gls"]
    if not os.path.isdir(os.path.join(OUT, "ninpol")):
        return False
    for n in need:
        d, b = os.path.split(n)
        dd = os.path.join(OUT, "ninpol", d)
        if not os.path.isdir(dd) or not any(f.startswith(b + ".") and f.endswith(".so") for f in os.listdir(dd)):
            return False
    return True


def build(force=False, verbose=True):
    if is_built() and not force:
        return OUT
    if not os.path.isdir(os.path.join(REF_SRC, "ninpol")):
        raise RuntimeError(f"reference sources not found at {REF_SRC}; cannot build oracle/_ref here")
    tmp = tempfile.mkdtemp(prefix="ninpol_ref_build_")
    try:
        shutil.copytree(os.path.join(REF_SRC, "ninpol"), os.path.join(tmp, "ninpol"))
        with open(os.path.join(tmp, "setup_ref.py"), "w") as f:
            f.write(_SETUP)
        env = dict(os.environ)
        # the image's default $CC (/opt/gcc) has no libgomp spec; the distro gcc does
        env["CC"] = "/usr/bin/gcc"
        env["CXX"] = "/usr/bin/g++"
        r = subprocess.run([sys.executable, "setup_ref.py", "build_ext", "--inplace"], cwd=tmp, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference build failed:\n" + r.stdout[-4000:])
        if os.path.isdir(OUT):
            shutil.rmtree(OUT)
        for root, _dirs, files in os.walk(os.path.join(tmp, "ninpol")):
            rel = os.path.relpath(root, tmp)
            for fn in files:
                keep = fn.endswith(".so") or fn.endswith(".py") or fn == "point_ordering.yaml"
                if keep:
                    os.makedirs(os.path.join(OUT, rel), exist_ok=True)
                    shutil.copy2(os.path.join(root, fn), os.path.join(OUT, rel, fn))
        if verbose:
            print(f"built reference into {OUT}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
