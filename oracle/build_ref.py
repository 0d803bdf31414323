"""oracle/build_ref.py -- TEST INFRASTRUCTURE.

Builds the parts of the REFERENCE ITSELF that can run in the authoring container, from the
sources where they lie under /root/reference, into oracle/_ref/ (git-ignored, travels to the
GPU box with the snapshot like any other built artefact):

  * lib/utils/div.pyx  -> oracle/_ref/cython_div*.so    unmodified
  * lib/utils/nms.pyx  -> oracle/_ref/cython_nms*.so    with the 2-token NumPy-2 shim
                          (np.int_t -> np.intp_t, dtype=np.int -> np.intp; semantics unchanged:
                          argsort returns intp)
  * lib/utils/bbox.pyx -> oracle/_ref/cython_bbox*.so   unmodified (recall parity only)
  * lib/detect/{test,tune,config}.py, lib/utils/{blob,timer}.py -> oracle/_ref/pyref/  mechanical
    py2 -> py3 text conversion (print statement, xrange, iteritems, has_key, cPickle, tabs),
    used ONLY by oracle/gen_golden.py in this container (and, when present, by bench.py's reference arm).
  * caffe-fast-rcnn/src/caffe/layers/{roi_pooling,grn,sigmoid,softmax,relu,inner_product,pooling}_layer.cpp
    -> oracle/_ref/libcaffe_layers_ref.so   UNMODIFIED sources, #included by oracle/ref_caffe_wrap.cpp and
    compiled against the stand-in framework headers of oracle/caffe_shim (Caffe itself cannot be built here);
    bindings in oracle/ref_caffe.py.

Nothing is copied into tracked files.  No-op (returns False) when /root/reference is absent,
e.g. on the GPU box, which uses the prebuilt .so files.
"""
from __future__ import annotations

import glob
import os
import re
import shutil
import subprocess
import sys
import tempfile

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

_SETUP = r"""
from setuptools import setup, Extension
from Cython.Build import cythonize
import numpy as np
exts = [Extension(n, [n + ".pyx"], include_dirs=[np.get_include()],
                  define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")],
                  extra_compile_args=["-O2", "-ffp-contract=off", "-w"])
        for n in ("cython_div", "cython_nms", "cython_bbox")]
setup(ext_modules=cythonize(exts, language_level=2, quiet=True))
"""


def _build_cython():
    with tempfile.TemporaryDirectory() as tmp:
        for name in ("div", "nms", "bbox"):
            src = open(os.path.join(REF, "lib", "utils", name + ".pyx")).read()
            if name == "nms":
                src = src.replace("np.int_t", "np.intp_t").replace("dtype=np.int)", "dtype=np.intp)")
            open(os.path.join(tmp, "cython_%s.pyx" % name), "w").write(src)
        open(os.path.join(tmp, "setup.py"), "w").write(_SETUP)
        subprocess.check_call([sys.executable, "setup.py", "-q", "build_ext", "--inplace"], cwd=tmp,
                              stdout=subprocess.DEVNULL)
        for so in glob.glob(os.path.join(tmp, "cython_*.so")):
            shutil.copy(so, OUT)


def _py2to3(text: str) -> str:
    text = text.replace("\t", "        ")
    text = text.replace("xrange(", "range(").replace(".iteritems()", ".items()")
    text = text.replace("import cPickle", "import pickle as cPickle")
    text = re.sub(r"(\w+)\.has_key\((\w+)\)", r"(\2 in \1)", text)
    text = text.replace("yaml.load(f)", "yaml.safe_load(f)")
    # print statements (possibly continued with a backslash) -> print(...)
    out, lines, i = [], text.split("\n"), 0
    while i < len(lines):
        ln = lines[i]
        m = re.match(r"^(\s*)print\s+(?!\()(.*)$", ln) or re.match(r"^(\s*)print\s+(\(.*\)\s*%.*|'.*)$", ln)
        if m:
            indent, body = m.group(1), m.group(2)
            while body.rstrip().endswith("\\"):
                i += 1
                body = body.rstrip()[:-1] + " " + lines[i].strip()
            # balance parentheses over following lines
            while body.count("(") > body.count(")"):
                i += 1
                body += " " + lines[i].strip()
            out.append("%sprint(%s)" % (indent, body))
        else:
            out.append(ln)
        i += 1
    return "\n".join(out)


def _convert_python():
    dst = os.path.join(OUT, "pyref")
    for sub in ("detect", "utils"):
        os.makedirs(os.path.join(dst, sub), exist_ok=True)
        open(os.path.join(dst, sub, "__init__.py"), "w").close()
    for rel in ("detect/test.py", "detect/tune.py", "detect/config.py", "utils/blob.py", "utils/timer.py"):
        src = _py2to3(open(os.path.join(REF, "lib", rel)).read())
        if rel == "detect/test.py":
            # two more edits that keep the Python-2 / NumPy-1 meaning: integer division and the
            # list-vs-array comparison `dets == []` (elementwise under NumPy 2)
            src = src.replace("max_per_set = 800 / (imdb.num_classes - 1)", "max_per_set = 800 // (imdb.num_classes - 1)")
            src = src.replace("if dets == []:", "if isinstance(dets, list) and dets == []:")
        open(os.path.join(dst, rel), "w").write(src)


def _build_caffe_layers():
    layers = os.path.join(REF, "caffe-fast-rcnn", "src", "caffe", "layers")
    so = os.path.join(OUT, "libcaffe_layers_ref.so")
    deps = [os.path.join(HERE, "ref_caffe_wrap.cpp"), os.path.join(HERE, "caffe_shim", "caffe", "shim.hpp")]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
        return
    # -ffp-contract=off: one rounding per float operation, like the reference's own (pre-FMA-default) builds
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                           "-I", os.path.join(HERE, "caffe_shim"), "-I", layers, deps[0], "-o", so])


def build() -> bool:
    if not os.path.isdir(os.path.join(REF, "lib", "utils")):
        return False
    os.makedirs(OUT, exist_ok=True)
    if not have_ref_cython():
        _build_cython()
    _convert_python()
    _build_caffe_layers()
    return True


def import_ref_cython():
    """Import the compiled reference Cython modules (needs the NumPy-1 aliases they use at
    module level: `DTYPE = np.float`, lib/utils/div.pyx:12)."""
    import importlib
    import numpy as np
    for alias, real in (("float", np.float64), ("int", np.intp), ("bool", np.bool_)):
        if not hasattr(np, alias):
            setattr(np, alias, real)
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    return (importlib.import_module("cython_div"), importlib.import_module("cython_nms"),
            importlib.import_module("cython_bbox"))


def load_pyref():
    """Import the converted reference modules of oracle/_ref/pyref (detect.test, detect.config) with the compiled
    reference Cython modules behind them and stand-ins for the absent `easydict` / `caffe` packages.  Nothing is
    built here: returns None when oracle/_ref is not populated.  The pyref tree shadows top-level `detect` and
    `utils`; callers that also use aznet_b200.detect import that through its package name, so the two do not clash."""
    import types
    pyref = os.path.join(OUT, "pyref")
    if not (have_ref_cython() and os.path.exists(os.path.join(pyref, "detect", "test.py"))):
        return None
    div, nms, bbox = import_ref_cython()

    class EasyDict(dict):                      # 20-line stand-in for the absent `easydict`
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                self[k] = v

        def __setitem__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setitem__(k, v)

        __setattr__ = __setitem__

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

    ed = types.ModuleType("easydict")
    ed.EasyDict = EasyDict
    sys.modules["easydict"] = ed
    sys.modules.setdefault("caffe", types.ModuleType("caffe"))
    if pyref not in sys.path:
        sys.path.insert(0, pyref)
    import utils  # noqa  (pyref/utils)
    sys.modules["utils.cython_nms"] = nms
    sys.modules["utils.cython_bbox"] = bbox
    sys.modules["utils.cython_div"] = div
    utils.cython_nms, utils.cython_bbox, utils.cython_div = nms, bbox, div
    import detect.test as rtest
    import detect.config as rconfig
    return rtest, rconfig, div, nms


def have_ref_cython() -> bool:
    return bool(glob.glob(os.path.join(OUT, "cython_div*.so"))) and bool(glob.glob(os.path.join(OUT, "cython_nms*.so")))


if __name__ == "__main__":
    print("built" if build() else "reference not present; nothing built")
