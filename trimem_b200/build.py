"""In-tree build of the native parts (no JIT cache: the .so files travel with the tree).

  trimem_b200/lib/libtrimem_b200.so   CUDA kernels + C ABI   (nvcc, sm_100a only)
  trimem_b200/core*.so                pybind11 host module   (g++, links the above)
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libtrimem_b200.so")
CORE = os.path.join(HERE, "core" + sysconfig.get_config_var("EXT_SUFFIX"))

CU_SOURCES = ["tm_api.cu", "tm_layout.cu", "tm_rings.cu", "tm_energy.cu", "tm_nlist.cu", "tm_flips.cu", "tm_geometry.cu",
              "tm_integrate.cu", "tm_part.cu"]
CU_HEADERS = ["tm_common.cuh", "tm_neigh.cuh", "tm_vertex.cuh", "tm_math.cuh"]
HOST_SOURCES = ["host/halfedge_mesh.cpp", "host/energy_manager.cpp", "host/mesh_io.cpp", "host/bindings.cpp"]
HOST_HEADERS = ["host/halfedge_mesh.h", "host/energy_manager.h"]
# the image exports CXX=/opt/gcc/bin/g++ whose OpenMP link is broken; use the system compiler
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--extended-lambda", "-Xcompiler", "-fPIC", "-ccbin", GXX]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd[:3]) + " ...")
    return r.stdout


def build_cuda(force=False, verbose=False, ptxas_info=False, defines=(), out=None):
    """defines/out: build a kernel variant (-DNAME=VALUE ...) into another .so (scripts/build_variant.py)."""
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in CU_HEADERS] + [os.path.join(ROOT, "include", "trimem_b200.h")]
    LIB = out or globals()["LIB"]
    if not force and not _newer(LIB, deps):
        return ""
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objs = []
    procs = []
    objdir = os.path.join(HERE, "build", os.path.basename(LIB) if out else "")
    os.makedirs(objdir, exist_ok=True)
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = ["nvcc"] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = ""
    for cmd, p in procs:
        out, _ = p.communicate()
        log += out
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed: " + " ".join(cmd[-3:]))
    log += _run(["nvcc", "-shared", "-o", LIB] + objs + ["-ccbin", GXX], verbose)
    return log


def build_host(force=False, verbose=False):
    import pybind11
    srcs = [os.path.join(CSRC, s) for s in HOST_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HOST_HEADERS] + [LIB, os.path.join(ROOT, "include", "trimem_b200.h")]
    if not force and not _newer(CORE, deps):
        return ""
    cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
           "-I" + sysconfig.get_paths()["include"], "-I" + pybind11.get_include()] + srcs + [
           "-L" + LIBDIR, "-ltrimem_b200", "-Wl,-rpath,$ORIGIN/lib", "-o", CORE]
    return _run(cmd, verbose)


def build_all(force=False, verbose=False):
    build_cuda(force, verbose)
    build_host(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
