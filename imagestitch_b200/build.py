"""Builds imagestitch_b200/libvfsms.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvfsms.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]
# translation unit -> extra flags.  surf.cu / match.cu keep CPU rounding (no FMA contraction): parity with oracle/.
UNITS = {
    "surf.cu": ["-fmad=false"],
    "match.cu": ["-fmad=false"],
    "match_tc.cu": ["-fmad=false"],
    "phase.cu": ["-fmad=false"],
    "blend.cu": ["-fmad=false"],
    "orb.cu": ["-fmad=false"],
    "enhance.cu": ["-fmad=false"],
    "jpeg.cu": [],
    "jpeg_enc.cu": [],
    "capi.cu": [],
}
LIBS = ["-lcufft"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "vfsms.h"))
    objs, procs = [], []
    for unit, extra in UNITS.items():
        src = os.path.join(CSRC, unit)
        if not os.path.exists(src):
            continue
        obj = os.path.join(objdir, unit.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC] + ARCH + COMMON + extra + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((unit, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for unit, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed on %s" % unit)
    if force or procs or _stale(OUT, objs):
        uses_fft = os.path.exists(os.path.join(CSRC, "phase.cu"))
        cmd = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-ccbin", "/usr/bin/g++"]
        if uses_fft:
            cmd += LIBS + ["-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
