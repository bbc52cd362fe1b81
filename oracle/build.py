"""Build recipes for the oracle (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

* ``build_oracle()``  gcc-compiles oracle/ovdet_oracle.c (our C restatement) into
  oracle/_build/libovdet_oracle.so.
* ``build_ref()``     when /root/reference is present (the build container only),
  compiles the reference's own Cython clipper utils/box_intersection.pyx *from
  where it lies* into oracle/_ref/ (cython -> C -> gcc; outputs only under
  oracle/_ref/, which is git-ignored but travels to the GPU box).  No reference
  source is copied into the repository.  The reference's own build script
  (utils/cython_compile.py:10) hard-codes a numpy<2 include path, so it is not
  used; include dirs come from numpy.get_include().
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
REF_DIR = os.path.join(HERE, "_ref")
ORACLE_SO = os.path.join(BUILD_DIR, "libovdet_oracle.so")
REFERENCE_ROOT = os.environ.get("OVDET_REFERENCE_ROOT", "/root/reference")
REF_PYX = os.path.join(REFERENCE_ROOT, "utils", "box_intersection.pyx")


def _newer(src, dst):
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build_oracle(force=False):
    src = os.path.join(HERE, "ovdet_oracle.c")
    if force or _newer(src, ORACLE_SO):
        os.makedirs(BUILD_DIR, exist_ok=True)
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
               "-o", ORACLE_SO, src, "-lm"]
        subprocess.check_call(cmd)
    return ORACLE_SO


def ref_so_path():
    suffix = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    return os.path.join(REF_DIR, "box_intersection" + suffix)


def build_ref(force=False):
    """Compile the reference Cython extension into oracle/_ref/. Returns the .so
    path, or None when the reference checkout is absent (e.g. on the GPU box,
    which uses the prebuilt file that travelled with the snapshot)."""
    so = ref_so_path()
    if not os.path.exists(REF_PYX):
        return so if os.path.exists(so) else None
    if not force and os.path.exists(so) and not _newer(REF_PYX, so):
        return so
    import numpy as np
    os.makedirs(REF_DIR, exist_ok=True)
    c_file = os.path.join(REF_DIR, "box_intersection.c")
    subprocess.check_call([sys.executable, "-m", "cython", "-3", REF_PYX, "-o", c_file])
    inc = sysconfig.get_paths()["include"]
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-fwrapv", "-fno-strict-aliasing",
           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
           "-I", inc, "-I", np.get_include(), "-o", so, c_file]
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    print("oracle:", build_oracle(force="--force" in sys.argv))
    print("ref   :", build_ref(force="--force" in sys.argv))
