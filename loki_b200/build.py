"""Build the sm_100a shared library `loki_b200/libloki_b200.so` with nvcc (in-tree, no JIT cache).

lk_kernels.cu is compiled twice: the production arithmetic (FMA contraction, namespace lkfast) and the
strict arithmetic (-fmad=false, reference operation order, namespace lkstrict).  The oracle under
oracle/ is built separately (it is test infrastructure, never linked here).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libloki_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "loki_b200.h"), os.path.join(HERE, "..", "include", "loki_b200_host.h"),
        os.path.abspath(__file__)]
    if not force and _newer(OUT, srcs):
        return OUT
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    jobs = [
        (["-DLK_STRICT=0"], "lk_kernels.cu", "lk_kernels_fast.o"),
        (["-DLK_STRICT=1", "-fmad=false"], "lk_kernels.cu", "lk_kernels_strict.o"),
        ([], "lk_capi.cu", "lk_capi.o"),
        ([], "lk_fft.cu", "lk_fft.o"),
        ([], "lk_diag.cu", "lk_diag.o"),
        (["-fmad=false"], "lk_bcs.cu", "lk_bcs.o"),
        (["-fmad=false"], "lk_f77.cu", "lk_f77.o"),
        (["-fmad=false"], "lk_flux.cu", "lk_flux.o"),
        (["-fmad=false"], "lk_coll.cu", "lk_coll.o"),
        ([], "lk_host.cu", "lk_host.o"),
    ]
    procs = []
    objs = []
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [
        os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))] + [
        os.path.abspath(__file__)]
    for flags, src, obj in jobs:
        if not os.path.exists(os.path.join(CSRC, src)):
            continue
        o = os.path.join(bdir, obj)
        objs.append(o)
        deps = [h for h in hdrs if src == "lk_host.cu" or not h.endswith("loki_b200_host.h")]
        if src != "lk_f77.cu":  # the Fortran-ABI prototypes only concern their own translation unit
            deps = [h for h in deps if not h.endswith("loki_b200_f77.h")]
        if "-DLK_STRICT=1" in flags or src != "lk_kernels.cu":  # the pipelined stage kernel: production build of lk_kernels.cu only
            deps = [h for h in deps if not h.endswith("lk_pipe.cuh")]
        if src != "lk_coll.cu":  # the collision operator's per-cell header
            deps = [h for h in deps if not h.endswith("lk_coll.cuh")]
        if not force and not verbose and _newer(o, [os.path.join(CSRC, src)] + deps):
            continue  # this object is current: only changed translation units are recompiled
        cmd = [nvcc] + ARCH + COMMON + extra + flags + ["-c", os.path.join(CSRC, src), "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs
    subprocess.check_call(cmd)
    return OUT


def build_variant(name, defines):
    """development aid: a second library `libloki_b200_<name>.so` whose production kernels are compiled
    with extra -D flags (A/B experiments on one GPU box; LOKI_B200_LIB selects it at load time).  The other
    objects are shared with the main build."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    build()
    bdir = os.path.join(HERE, "build")
    o = os.path.join(bdir, "lk_kernels_fast_%s.o" % name)
    cmd = [nvcc] + ARCH + COMMON + ["-DLK_STRICT=0"] + ["-D" + d for d in defines] + ["-c", os.path.join(CSRC, "lk_kernels.cu"), "-o", o]
    subprocess.check_call(cmd)
    out = os.path.join(HERE, "libloki_b200_%s.so" % name)
    objs = [o] + [os.path.join(bdir, f) for f in ("lk_kernels_strict.o", "lk_capi.o", "lk_fft.o", "lk_diag.o", "lk_bcs.o", "lk_f77.o", "lk_flux.o", "lk_coll.o", "lk_host.o")]
    subprocess.check_call([nvcc] + ARCH + ["-shared", "-o", out] + objs)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
