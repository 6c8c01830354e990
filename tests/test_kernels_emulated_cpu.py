"""CPU: the CUDA kernels themselves, executed without a GPU.

tests/cuda_emu/ compiles the product's .cu sources (imagestitch_b200/csrc, minus the tcgen05 matcher) with g++ on a small
CUDA-on-CPU execution model (fibers per thread, lock-step warp collectives, block barriers, textures, a DFT in place of
cuFFT).  The `gpu`-marked parity tests of this directory then run in a subprocess against that library (VFSMS_EMU=1, see
conftest.py): same C ABI, same host code, same oracle comparisons.  This checks kernel LOGIC -- indexing, reductions, work
distribution, exact arithmetic order -- in every CPU run, including the code written while no GPU was available (kernel
variants behind vfsms_set_option, vfsms_mosaic_band_host, vfsms_overlap_sums_host).  It says nothing about timing, memory-model
races or the tensor-core path; the real `-m gpu` run on the B200 remains the parity gate.

Default selection: about 1.5 minutes.  VFSMS_EMU_FULL=1 runs every gpu test the emulation covers (about 10 minutes).
"""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "cuda_emu")

pytestmark = pytest.mark.skipif(os.environ.get("VFSMS_EMU") == "1", reason="already inside the emulated run")

# (selection, -k expression, minimum number of tests that must have passed)
FAST = [
    (["tests/test_gpu_blend.py", "tests/test_gpu_zz_bands.py", "tests/test_gpu_phase.py", "tests/test_gpu_orb.py"], None, 20),
    (["tests/test_gpu_jpeg.py", "tests/test_gpu_zz_jpeg_encode.py"], "not full_size", 35),
    (["tests/test_gpu_surf.py"], "not real_micrograph", 9),
    (["tests/test_gpu_zz_phase_wrap.py", "tests/test_gpu_variants.py", "tests/test_gpu_zz_colour_stack.py", "tests/test_gpu_zz_entropy_device.py",
      "tests/test_gpu_zz_sequence.py"],
     "overlap_sums or sort_per_image or large_windows_first or borders_and_giants or colour_twin or (entropy_device and not large) or result_file_identical", 9),
]


@pytest.fixture(scope="module")
def emu_lib():
    sys.path.insert(0, EMU_DIR)
    try:
        import build_emu
        return build_emu.build()
    finally:
        sys.path.remove(EMU_DIR)


def _start(selection, kexpr, extra_env=None):
    """pytest on the emulated library in a subprocess (output to a temporary file: no pipe to fill up while nobody reads)"""
    import tempfile
    env = dict(os.environ, VFSMS_EMU="1", VFSMS_EXPERIMENTAL="1", PYTHONPATH=ROOT)
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"] + selection
    if kexpr:
        cmd += ["-k", kexpr]
    log = tempfile.TemporaryFile()
    return subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=log, stderr=subprocess.STDOUT), log


def _finish(started, timeout=3000):
    proc, log = started
    try:
        proc.wait(timeout=timeout)
    except subprocess.TimeoutExpired:
        proc.kill()
        proc.wait()
    log.seek(0)
    out = log.read().decode(errors="replace")
    log.close()
    m = re.search(r"(\d+) passed", out)
    return proc.returncode, int(m.group(1)) if m else 0, out


def _run(selection, kexpr, extra_env=None):
    return _finish(_start(selection, kexpr, extra_env))


@pytest.fixture(scope="module")
def emu_runs(emu_lib):
    """All selections start together (one pytest subprocess each, the library is built once before): the wall time of this
    module is the slowest selection, not their sum."""
    runs = {("fast", i): _start(sel, kexpr) for i, (sel, kexpr, _) in enumerate(FAST)}
    for limit in ("450", "250"):
        runs[("tex", limit)] = _start(["tests/test_gpu_variants.py"], "stacked_texture_batches", {"VFSMS_EMU_TEX_ROWS": limit})
    yield runs
    for proc, log in runs.values():
        if proc.poll() is None:
            proc.kill()
            proc.wait()


def test_emulated_library_exports_the_c_abi(emu_lib):
    """the emulated build links without libcudart / libcufft and exports every symbol of include/vfsms.h"""
    from imagestitch_b200 import _lib
    L = ctypes.CDLL(emu_lib)
    for name in _lib.EXPORTS:
        getattr(L, name)
    ldd = subprocess.run(["ldd", emu_lib], stdout=subprocess.PIPE).stdout.decode()
    assert "libcudart" not in ldd and "libcufft" not in ldd and "libcuda" not in ldd
    assert _lib.SO_PATH.endswith(os.path.join("imagestitch_b200", "libvfsms.so")), "the product must never point at the emulation"


@pytest.mark.parametrize("case", range(len(FAST)))
def test_gpu_parity_tests_on_emulated_kernels(emu_runs, case):
    selection, kexpr, at_least = FAST[case]
    rc, passed, out = _finish(emu_runs[("fast", case)])
    assert rc == 0, out[-6000:]
    assert passed >= at_least, out[-2000:]


def test_stacked_texture_groups_on_emulated_kernels(emu_runs):
    """describe mode 2 with a 450-row (groups of 4 images) and a 250-row (groups of 2) texture height limit"""
    for limit in ("450", "250"):
        rc, passed, out = _finish(emu_runs[("tex", limit)])
        assert rc == 0 and passed == 1, out[-6000:]


def _tsan_runtime():
    for cxx in ("/usr/bin/g++", "g++"):
        try:
            p = subprocess.run([cxx, "-print-file-name=libtsan.so"], stdout=subprocess.PIPE).stdout.decode().strip()
        except OSError:
            continue
        if os.path.isabs(p) and os.path.exists(p):
            return p
    return None


def test_racecheck_controls_and_kernels():
    """ThreadSanitizer build of the emulation: CUDA threads are TSan fibers ordered only by barriers / warp collectives.  The controls
    prove the detector works (a missing __syncthreads and a missing __syncwarp are reported, the synchronised version is not), then
    small inputs go through every kernel family: no intra-block race."""
    tsan = _tsan_runtime()
    if tsan is None:
        pytest.skip("no libtsan.so for the system g++")
    env = dict(os.environ, VFSMS_EMU_SANITIZE="thread", PYTHONPATH=ROOT)
    builder = os.path.join(EMU_DIR, "build_emu.py")
    exe = subprocess.run([sys.executable, builder, "--selftest"], env=env, stdout=subprocess.PIPE, check=True).stdout.decode().split()[-1]
    subprocess.run([sys.executable, builder], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    opts = "suppressions=%s:exitcode=0:report_signal_unsafe=0" % os.path.join(EMU_DIR, "tsan.supp")
    run_env = dict(env, TSAN_OPTIONS=opts)
    for what, racy in (("clean", False), ("racy", True), ("racy_warp", True)):
        p = subprocess.run([exe, what], env=run_env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        assert p.returncode == 0 and b"finished" in p.stdout, (what, p.stdout, p.stderr[-2000:])
        assert (p.stderr.count(b"WARNING: ThreadSanitizer: data race") > 0) == racy, (what, p.stderr[-3000:])
    p = subprocess.run([sys.executable, os.path.join(EMU_DIR, "racecheck.py")], env=dict(run_env, LD_PRELOAD=tsan), cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=1500)
    assert p.returncode == 0 and b"racecheck done" in p.stdout, (p.stdout[-2000:], p.stderr[-3000:])
    assert p.stderr.count(b"WARNING: ThreadSanitizer") == 0, p.stderr[-6000:].decode(errors="replace")


@pytest.mark.skipif(os.environ.get("VFSMS_EMU_FULL") != "1", reason="set VFSMS_EMU_FULL=1 (about 10 minutes)")
def test_all_emulable_gpu_tests(emu_lib):
    rc, passed, out = _run(["tests"], None)
    assert rc == 0, out[-6000:]
    assert passed >= 100, out[-2000:]
