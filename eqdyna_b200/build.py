"""Build recipes for the native parts (in-tree, so the .so files travel with the
gpurun snapshot).

  libeqdyna_host.so  -- stand-in Fortran host (g++, no CUDA)
  libeqdyna_b200.so  -- the CUDA step library, sm_100a (nvcc)
  eqdyna_host        -- standalone driver executable (links both)
  oracle/_build/liboracle.so -- CPU oracle (test infrastructure; g++)

`python -m eqdyna_b200.build [host|cuda|oracle|exe|all]`
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "eqdyna_b200")
CSRC = os.path.join(PKG, "csrc")
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(PKG, "lib")
ORACLE = os.path.join(ROOT, "oracle")

HOST_SRCS = ["eqh_io.cpp", "eqh_mesh.cpp", "eqh_mass.cpp", "eqh_api.cpp"]
CUDA_SRCS = ["eqd_api.cu", "eqd_kernels.cu", "eqd_tiles.cu", "eqd_ops.cu", "eqd_march.cu"]
# the operator precompute must round like the reference build (no FMA contraction)
CUDA_FILE_FLAGS = {"eqd_ops.cu": ["--fmad=false"]}

# the reference's ubuntu build is -O3 without -march / fast-math: no FMA contraction
HOST_FLAGS = ["-O2", "-ffp-contract=off", "-fPIC", "-std=c++17", "-Wall", "-Wno-unused-variable"]
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def _headers():
    hs = [os.path.join(INC, f) for f in os.listdir(INC)]
    for d in ("host", "cuda"):
        p = os.path.join(CSRC, d)
        if os.path.isdir(p):
            hs += [os.path.join(p, f) for f in os.listdir(p) if f.endswith((".h", ".cuh"))]
    return hs


def host_lib_path():
    return os.path.join(LIBDIR, "libeqdyna_host.so")


def cuda_lib_path():
    return os.path.join(LIBDIR, "libeqdyna_b200.so")


def oracle_lib_path():
    return os.path.join(ORACLE, "_build", "liboracle.so")


def build_host(force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, "host", s) for s in HOST_SRCS]
    out = host_lib_path()
    if force or _newer(out, srcs + _headers()):
        _run(["g++"] + HOST_FLAGS + ["-shared", "-I", INC, "-o", out] + srcs)
    return out


def build_oracle(force=False):
    os.makedirs(os.path.join(ORACLE, "_build"), exist_ok=True)
    srcs = [os.path.join(ORACLE, "step_oracle.cpp")]
    out = oracle_lib_path()
    if force or _newer(out, srcs + _headers()):
        _run(["g++"] + HOST_FLAGS + ["-fopenmp", "-shared", "-I", INC, "-o", out] + srcs)
    return out


def nccl_flags():
    """NCCL is dlopen'ed at run time (eqd_api.cu): headers from the torch-bundled
    package when present (else /usr/include/nccl.h); the bundled library's path is
    baked in as a fallback for processes that have not loaded libnccl.so.2 yet."""
    inc, libs = [], []
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            base = list(spec.submodule_search_locations)[0]
            if os.path.exists(os.path.join(base, "include", "nccl.h")):
                inc = ["-I", os.path.join(base, "include")]
            lib = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(lib):
                libs = ['-DEQD_NCCL_PATH="%s"' % lib]
    except Exception:
        pass
    return inc, libs + ["-ldl"]


def build_cuda(force=False, extra=()):
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, "cuda", s) for s in CUDA_SRCS]
    out = cuda_lib_path()
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    inc, libs = nccl_flags()
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    hdrs = _headers()
    objs, jobs = [], []
    for src in srcs:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            cmd = [nvcc, "-O3", "-std=c++17", "-lineinfo"] + NVCC_ARCH + ["-Xcompiler", "-fPIC", "-I", INC] + inc + [
                f for f in libs if f.startswith("-D")] + CUDA_FILE_FLAGS.get(os.path.basename(src), []) + list(extra) + ["-c", src, "-o", obj]
            jobs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
            jobs[-1].cmdline = cmd
    for j in jobs:   # the translation units compile side by side
        o, _ = j.communicate()
        if j.returncode != 0:
            raise RuntimeError("build failed: %s\n%s" % (" ".join(j.cmdline), o))
    if force or jobs or _newer(out, objs):
        _run([nvcc, "-shared"] + NVCC_ARCH + ["-o", out] + objs + [f for f in libs if not f.startswith("-D")])
    return out


def build_exe(force=False):
    host = build_host(force)
    cuda = build_cuda(force)
    src = os.path.join(CSRC, "host", "eqdyna_host_main.cpp")
    out = os.path.join(PKG, "bin", "eqdyna_host")
    if not os.path.exists(src):
        return None
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if force or _newer(out, [src, host, cuda] + _headers()):
        _run(["g++"] + HOST_FLAGS + ["-I", INC, "-o", out, src, host, cuda,
                                      "-Wl,-rpath," + LIBDIR])
    return out


def build_all(force=False):
    out = {"host": build_host(force), "cuda": build_cuda(force), "oracle": build_oracle(force)}
    exe = build_exe(force)   # standalone driver (eqdyna_host <case_dir>): links the two libraries above
    if exe:
        out["exe"] = exe
    return out


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    fn = {"host": build_host, "cuda": build_cuda, "oracle": build_oracle, "exe": build_exe, "all": build_all}[what]
    print(fn(force="--force" in sys.argv))
