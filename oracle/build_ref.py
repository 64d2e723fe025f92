#!/usr/bin/env python
"""Recipe: compile the reference's only native source, psoap/matrix_functions.pyx, into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  The source is read where it lies under /root/reference (never copied into the
repo); every output (generated C, object, shared library) goes to oracle/_ref/, which is git-ignored but
travels to the GPU box with the snapshot.  The reference's own build system (setup.py) is not run: this is
one `cython` call and one `gcc` call with the flags setup.py:30-31 implies (numpy include dir, -O2).

Usage: python oracle/build_ref.py [--force]
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/psoap/matrix_functions.pyx"
OUT_DIR = os.path.join(HERE, "_ref", "psoap")


def ref_so_path():
    return os.path.join(OUT_DIR, "matrix_functions" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False):
    so = ref_so_path()
    if not os.path.exists(REF_PYX):
        return so if os.path.exists(so) else None  # GPU box: use the prebuilt file if it travelled
    if os.path.exists(so) and not force and os.path.getmtime(so) >= os.path.getmtime(REF_PYX):
        return so
    import numpy as np
    os.makedirs(OUT_DIR, exist_ok=True)
    c_file = os.path.join(OUT_DIR, "matrix_functions.c")
    subprocess.check_call([sys.executable, "-m", "cython", "-3", "--module-name", "psoap.matrix_functions",
                           REF_PYX, "-o", c_file])
    inc = sysconfig.get_paths()["include"]
    subprocess.check_call(["gcc", "-shared", "-fPIC", "-O2", "-fwrapv", "-fno-strict-aliasing",
                           "-DNPY_NO_DEPRECATED_API=0", "-I", inc, "-I", np.get_include(), c_file, "-o", so])
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
