"""Build recipe of libvkrt_cuda.so (nvcc, sm_100a only, in-tree so the .so travels with gpurun).

The vkrt-f32 arithmetic contract needs -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
(csrc/vkrt_arith.cuh refuses to compile without -DVKRT_FMAD_OFF, which is set next to them).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvkrt_cuda.so")
OBJ = os.path.join(HERE, "..", "build")
SOURCES = ["vkrt_render.cu", "vkrt_wavefront.cu", "vkrt_bvh.cu", "vkrt_api.cu", "vkrt_exchange.cu", "vkrt_micro.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-DVKRT_FMAD_OFF",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def _newer(src, dst):
    return not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: tuning builds (e.g. defines=["VKRT_REFILL=24"], out="vk-renderer_b200/libvkrt_r24.so",
    selected at run time with the VKRT_LIB environment variable); the default build has neither."""
    global OBJ
    nvcc = os.environ.get("NVCC", "nvcc")
    out = out or OUT
    if defines:
        OBJ = os.path.join(HERE, "..", "build", "_".join(d.replace("=", "") for d in defines))
        force = True
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "vkrt.h"))
    hdr_time = max(os.path.getmtime(h) for h in headers)
    objs, rebuilt = [], False
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if force or _newer(src, obj) or hdr_time > os.path.getmtime(obj):
            cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            rebuilt = True
    for s, p in procs:
        log, _ = p.communicate()
        if verbose and log:
            sys.stderr.write(log)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (s, log))
    if rebuilt or not os.path.exists(out):
        cmd = [nvcc, "-shared", "-cudart", "static", "-Wno-deprecated-gpu-targets", "-o", out] + objs
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    return out


HOST_SOURCES = ["Camera.cpp", "GraphicsDevice_cuda.cpp", "headless_main.cpp"]
HOST_OUT = os.path.join(HERE, "vkrt_headless")


def build_host(force=False):
    """g++ build of the C++ host side (csrc/host): the GraphicsDevice drop-in + the headless frame loop, linked
    against libvkrt_cuda.so (rpath $ORIGIN).  -ffp-contract=off: Camera.cpp must round like the reference's build."""
    srcs = [os.path.join(CSRC, "host", f) for f in HOST_SOURCES]
    deps = srcs + [os.path.join(CSRC, "host", "GraphicsDevice.h"), os.path.join(HERE, "..", "include", "vkrt.h"), OUT]
    if not force and os.path.exists(HOST_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_OUT) for d in deps):
        return HOST_OUT
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Wextra",
           "-I" + os.path.join(HERE, "..", "include"), "-I" + os.path.join(CSRC, "host")] + srcs + \
          ["-L" + HERE, "-lvkrt_cuda", "-Wl,-rpath,$ORIGIN", "-o", HOST_OUT]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build failed:\n" + r.stdout)
    return HOST_OUT


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
    if not defs and not outs:
        print(build_host(force="--force" in sys.argv))
