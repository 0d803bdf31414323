#!/usr/bin/env python
"""Previous-build A/B: link a second copy of the library in which ONE source file comes from another git revision (or is
compiled with extra flags), so that a kernel change can be measured against the previous BUILD on the same box -- an
A/B between two paths of one binary says nothing about what the new code did to the old path (the ROI-pool regression of
round 2, profiles/r2_pool_regression_ab.txt).

    python tools/ab_build.py --name old --file roi_pool.cu --ref ee33f6f          # that file as of the revision
    python tools/ab_build.py --name u1  --file roi_pool.cu --flags=-DAZN_POOL_PW_UNROLL=1

writes aznet_b200/build/ab/lib_<name>.so (git-ignored, travels with the gpurun snapshot).  On the box:

    bash tools/ab_run.sh <name> python tools/microbench.py --only roi_pool     # runs the command with that library swapped in

The other objects are the current build's (aznet_b200/build/*.o): the revision must export the same C ABI for that file."""
import argparse
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aznet_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", required=True)
    ap.add_argument("--file", required=True, help="a source under aznet_b200/csrc")
    ap.add_argument("--ref", default="", help="git revision to take the file from (default: the working tree)")
    ap.add_argument("--flags", default="", help="extra nvcc flags for that file")
    args = ap.parse_args()
    _lib.build()                                             # the current objects
    csrc = os.path.join(ROOT, "aznet_b200", "csrc")
    build = os.path.join(ROOT, "aznet_b200", "build")
    out_dir = os.path.join(build, "ab")
    os.makedirs(out_dir, exist_ok=True)
    stem = os.path.splitext(args.file)[0]
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(csrc, args.file)
        if args.ref:
            text = subprocess.check_output(["git", "-C", ROOT, "show", "%s:aznet_b200/csrc/%s" % (args.ref, args.file)])
            src = os.path.join(csrc, "_ab_%s_%s" % (args.name, args.file))      # beside the headers it includes
            with open(src, "wb") as f:
                f.write(text)
        obj = os.path.join(tmp, stem + ".o")
        try:
            cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + [f for f in _lib.NVCC_FLAGS if f != "-shared"] + args.flags.split() + ["-c", src, "-o", obj]
            subprocess.check_call(cmd)
        finally:
            if args.ref and os.path.exists(src):
                os.remove(src)
        others = [os.path.join(build, o) for o in sorted(os.listdir(build)) if o.endswith(".o") and o != stem + ".o"]
        so = os.path.join(out_dir, "lib_%s.so" % args.name)
        subprocess.check_call([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-Xcompiler", "-pthread", "-o", so] + others + [obj])
    print("built", so)


if __name__ == "__main__":
    main()
