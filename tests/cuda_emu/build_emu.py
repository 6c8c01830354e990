"""TEST INFRASTRUCTURE ONLY.  Builds tests/cuda_emu/_build/libvfsms_emu.so: the .cu sources of imagestitch_b200/csrc compiled by g++
against the CUDA-on-CPU execution model of emu.h / emu.cpp, so that kernel logic can be exercised without a GPU.

The only source transformation is textual and local:
  * `kernel<T...><<<grid, block, smem, stream>>>(args)`  ->  `emu::launch_cfg([&]() { kernel<T...>(args); }, grid, block, smem, stream)`
  * `extern __shared__ T name[];`                          ->  `T *name = (T *)emu::dyn_smem;`
Everything else (threadIdx, __shared__, warp collectives, atomics, textures, the CUDA runtime calls) is supplied by the
force-included emu.h.  match_tc.cu (tcgen05 / TMA inline PTX) is not compiled; emu.cpp holds a stub that reports so.
Nothing in the product imports this module or loads the library it builds.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "imagestitch_b200", "csrc")
# VFSMS_EMU_SANITIZE=1: AddressSanitizer build in its own directory (run python with LD_PRELOAD=$(g++ -print-file-name=libasan.so))
# VFSMS_EMU_SANITIZE=thread: ThreadSanitizer build (LD_PRELOAD libtsan.so, TSAN_OPTIONS=suppressions=tests/cuda_emu/tsan.supp): racecheck
SANITIZE = os.environ.get("VFSMS_EMU_SANITIZE") in ("1", "thread")
TSAN = os.environ.get("VFSMS_EMU_SANITIZE") == "thread"
OUT_DIR = os.path.join(HERE, "_build", "tsan" if TSAN else "asan") if SANITIZE else os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libvfsms_emu.so")
UNITS = ["surf.cu", "match.cu", "phase.cu", "blend.cu", "orb.cu", "enhance.cu", "jpeg.cu", "jpeg_enc.cu", "capi.cu"]
CXX = os.environ.get("VFSMS_EMU_CXX") or ("/usr/bin/g++" if SANITIZE and os.path.exists("/usr/bin/g++") else os.environ.get("CXX", "g++"))
# -ffp-contract=off mirrors nvcc -fmad=false (parity with oracle/ depends on unfused arithmetic)
FLAGS = ["-O1", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-w", "-pthread",
         "-I/usr/local/cuda/include", "-I" + CSRC, "-include", os.path.join(HERE, "emu.h")]
if SANITIZE:
    FLAGS += ["-fsanitize=thread" if TSAN else "-fsanitize=address", "-fno-omit-frame-pointer"]


def _match_back(text, pos):
    """text[pos-1] == '>' : index of the matching '<' scanning backwards."""
    depth = 0
    i = pos - 1
    while i >= 0:
        c = text[i]
        if c == ">":
            depth += 1
        elif c == "<":
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template arguments before <<<")


def _match_fwd(text, pos):
    """text[pos] == '(' : index just after the matching ')'."""
    depth = 0
    i = pos
    while i < len(text):
        c = text[i]
        if c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced launch arguments")


def translate(text):
    out = []
    pos = 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            out.append(text[pos:])
            break
        # kernel expression: identifier [ <template args> ] directly before <<<
        e = k
        while e > 0 and text[e - 1] in " \t":
            e -= 1
        s = e
        if text[s - 1] == ">":
            s = _match_back(text, s)
            while s > 0 and text[s - 1] in " \t":
                s -= 1
        while s > 0 and (text[s - 1].isalnum() or text[s - 1] in "_:"):
            s -= 1
        kernel = text[s:e]
        c_end = text.find(">>>", k)
        cfg = text[k + 3:c_end]
        a = c_end + 3
        while text[a] in " \t\\\n":
            a += 1
        if text[a] != "(":
            raise ValueError("no argument list after >>> near: " + text[k - 40:k + 80])
        a_end = _match_fwd(text, a)
        args = text[a:a_end]
        out.append(text[pos:s])
        out.append("emu::launch_cfg([&]() { %s%s; }, %s)" % (kernel, args, cfg))
        pos = a_end
    text = "".join(out)
    text = re.sub(r"extern\s+__shared__\s+([A-Za-z_0-9]+)\s+([A-Za-z_0-9]+)\s*\[\s*\]\s*;",
                  r"\1 *\2 = (\1 *)emu::dyn_smem;", text)
    return text


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    common = [os.path.join(HERE, f) for f in ("emu.h", "build_emu.py")]
    common += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))]
    common.append(os.path.join(ROOT, "include", "vfsms.h"))
    jobs, objs = [], []
    # headers that launch kernels are translated too; the generated sources live in OUT_DIR, so their `#include "x.cuh"` finds the copy
    for name in os.listdir(CSRC):
        if name.endswith(".cuh"):
            with open(os.path.join(CSRC, name)) as f:
                text = f.read()
            if "<<<" in text:
                gen = os.path.join(OUT_DIR, name)
                body = '#line 1 "%s"\n' % os.path.join(CSRC, name) + translate(text)
                if not os.path.exists(gen) or open(gen).read() != body:
                    with open(gen, "w") as f:
                        f.write(body)
    for unit in UNITS:
        src = os.path.join(CSRC, unit)
        gen = os.path.join(OUT_DIR, unit.replace(".cu", "_emu.cpp"))
        obj = gen.replace(".cpp", ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + common):
            with open(src) as f:
                body = translate(f.read())
            with open(gen, "w") as f:
                f.write('#line 1 "%s"\n' % src + body)
            cmd = [CXX] + FLAGS + ["-c", gen, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            jobs.append((unit, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    emu_obj = os.path.join(OUT_DIR, "emu.o")
    objs.append(emu_obj)
    if force or _stale(emu_obj, [os.path.join(HERE, "emu.cpp")] + common):
        emu_flags = [f for f in FLAGS if f != "-fsanitize=thread"] + (["-DEMU_FORCE_TSAN"] if TSAN else [])
        cmd = [CXX] + emu_flags + ["-c", os.path.join(HERE, "emu.cpp"), "-o", emu_obj]
        jobs.append(("emu.cpp", subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for unit, p in jobs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode()[-20000:])
            raise RuntimeError("g++ failed on %s (emulated build)" % unit)
    if force or jobs or _stale(OUT, objs):
        cmd = [CXX, "-shared", "-o", OUT] + objs + ["-pthread"] + (["-fsanitize=thread" if TSAN else "-fsanitize=address"] if SANITIZE else ["-Wl,--no-undefined"])
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return OUT


def build_selftest():
    """selftest.cu (controls of the execution model and of the racecheck) as a standalone executable in OUT_DIR."""
    os.makedirs(OUT_DIR, exist_ok=True)
    gen = os.path.join(OUT_DIR, "selftest_emu.cpp")
    with open(os.path.join(HERE, "selftest.cu")) as f:
        body = translate(f.read())
    with open(gen, "w") as f:
        f.write(body)
    exe = os.path.join(OUT_DIR, "selftest")
    emu_o = os.path.join(OUT_DIR, "emu_standalone.o")
    emu_flags = [f for f in FLAGS if f != "-fsanitize=thread"] + (["-DEMU_FORCE_TSAN"] if TSAN else [])
    subprocess.check_call([CXX] + emu_flags + ["-DEMU_STANDALONE", "-c", os.path.join(HERE, "emu.cpp"), "-o", emu_o])
    subprocess.check_call([CXX] + FLAGS + [gen, emu_o, "-o", exe])
    return exe


if __name__ == "__main__":
    if "--selftest" in sys.argv:
        print(build_selftest())
    else:
        print(build(force="--force" in sys.argv, verbose=True))
