"""In-tree build of the native pieces (sm_100a only):

  edgefem_b200/lib/libedgefem_b200.so   CUDA kernels + C-ABI        (nvcc, edgefem_b200/csrc)
  edgefem_b200/lib/libedgefem.so        C++ host API (edgefem::)    (g++,  edgefem_b200/host)
  edgefem_b200/pyedgefem*.so            pybind11 module             (g++,  edgefem_b200/python)

Artefacts stay in the tree (git-ignored) so they travel to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(HERE, "lib")


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return r.stdout


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_cuda(verbose=False):
    out = _run(["make", "-j4"], cwd=os.path.join(HERE, "csrc"))
    if verbose:
        print(out)


def build_host(verbose=False):
    import pybind11

    os.makedirs(LIBDIR, exist_ok=True)
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "host")]
    host_srcs = [os.path.join(HERE, "host", f) for f in ("mesh_ports.cpp", "device_path.cpp", "port_modes_io.cpp", "post.cpp")]
    hdrs = []
    for d, _, fs in os.walk(os.path.join(ROOT, "include")):
        hdrs += [os.path.join(d, f) for f in fs]
    hdrs.append(os.path.join(HERE, "host", "host_internal.hpp"))
    libhost = os.path.join(LIBDIR, "libedgefem.so")
    common = ["g++", "-O2", "-std=c++17", "-fPIC", "-Wall", "-Wno-unused-function", "-fvisibility=default"]
    if not _newer(libhost, host_srcs + hdrs + [os.path.join(LIBDIR, "libedgefem_b200.so")]):
        _run(common + ["-shared", "-o", libhost] + host_srcs + inc + ["-L" + LIBDIR, "-ledgefem_b200", "-Wl,-rpath,$ORIGIN", "-pthread"])
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    mod = os.path.join(HERE, "pyedgefem" + ext)
    src = os.path.join(HERE, "python", "pyedgefem.cpp")
    if not _newer(mod, [src, libhost] + hdrs):
        _run(common + ["-shared", "-o", mod, src] + inc + ["-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"],
                                                          "-L" + LIBDIR, "-ledgefem", "-ledgefem_b200", "-Wl,-rpath,$ORIGIN/lib", "-pthread"])
    return mod


def build_all(verbose=False):
    build_cuda(verbose)
    return build_host(verbose)


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv))
