// emu.cpp -- TEST INFRASTRUCTURE ONLY (see emu.h): fiber scheduler of the CUDA-on-CPU execution model + the CUDA runtime / cuFFT
// entry points the host side of libvfsms calls.  Linked only into tests/cuda_emu/_build/libvfsms_emu.so.
#include "emu.h"
#include <cufft.h>
#include <sys/mman.h>
#include <complex>
#include <map>
#include <vector>

#ifndef EMU_STANDALONE
#include "../../imagestitch_b200/csrc/common.cuh"
#endif

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

// ---------------------------------------------------------------- fibers
// Minimal x86-64 SysV context switch: callee-saved registers on the old stack, swap rsp.  (ucontext's swapcontext makes a
// sigprocmask system call per switch; a warp shuffle is 64 switches.)
extern "C" void emu_switch(void **save_sp, void *new_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

// AddressSanitizer build (VFSMS_EMU_SANITIZE=1 in build_emu.py): device memory is malloc'ed and __shared__ arrays are statics, so
// out-of-bounds global / shared accesses of a kernel are reported like compute-sanitizer's memcheck would.  ASan has to be told
// about the stack switches.
#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/common_interface_defs.h>
#define EMU_ASAN 1
#else
#define EMU_ASAN 0
#endif

// ThreadSanitizer build (VFSMS_EMU_SANITIZE=thread): every CUDA thread of a block is a TSan fiber; switches create NO
// happens-before edge, only __syncthreads / warp collectives (release + acquire on a per-block / per-warp token) and the block and
// launch boundaries do.  Accesses of two threads of a block to the same shared or global word with no barrier between them are then
// reported although the emulation runs them one after the other -- compute-sanitizer's racecheck, for intra-block races.  Blocks are
// ordered one after the other (inter-block races are not modelled).
// emu.cpp itself is compiled WITHOUT -fsanitize=thread (-DEMU_FORCE_TSAN instead): the scheduler's own bookkeeping is then invisible to
// the detector, only the kernels' accesses are checked.
#if defined(__SANITIZE_THREAD__) || defined(EMU_FORCE_TSAN)
#include <sanitizer/tsan_interface.h>
#define EMU_TSAN 1
#else
#define EMU_TSAN 0
#endif

namespace emu {

enum { RUN = 0, AT_BLOCK = 1, AT_WARP = 2, DONE = 3 };
struct Fiber { void *sp; int state; uint3 tid; uint64_t slot; void *fake; };
static const void *sched_bottom = nullptr;
static size_t sched_size = 0;
#if EMU_TSAN
static void *tsan_fibers[1024];
static void *tsan_sched = nullptr;
// launch token: released by the scheduler before a block, acquired by every thread at its start; done token: released by every thread
// at its end, acquired by the scheduler after the block (two tokens: a thread's end must not order the next thread's start)
static char tsan_block_token, tsan_launch_token, tsan_done_token, tsan_warp_token[32];
#endif

static const size_t STACK_BYTES = 512 << 10;
static const int MAX_THREADS = 1024;
static char *stacks = nullptr;
static Fiber fibers[MAX_THREADS];
static void *sched_sp;
static int cur = -1, n_threads = 0;
static const std::function<void()> *cur_body = nullptr;
char *dyn_smem = nullptr;
static const size_t DYN_SMEM_BYTES = 256 << 10;

static void fiber_main()
{
#if EMU_ASAN
    __sanitizer_finish_switch_fiber(nullptr, &sched_bottom, &sched_size);
#endif
#if EMU_TSAN
    __tsan_acquire(&tsan_launch_token);       // everything before this block (host code, earlier blocks) happened before
#endif
    (*cur_body)();
#if EMU_TSAN
    __tsan_release(&tsan_done_token);
    __tsan_switch_to_fiber(tsan_sched, __tsan_switch_to_fiber_no_sync);
#endif
    fibers[cur].state = DONE;
    void *dummy;
#if EMU_ASAN
    __sanitizer_start_switch_fiber(nullptr, sched_bottom, sched_size);      // nullptr: this fiber's stack is abandoned
#endif
    emu_switch(&dummy, sched_sp);
    abort();
}

static void init_once()
{
    if (stacks) return;
    stacks = (char *)mmap(nullptr, STACK_BYTES * MAX_THREADS, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (stacks == MAP_FAILED) { perror("emu: mmap"); abort(); }
    for (int i = 0; i < MAX_THREADS; i++) mprotect(stacks + (size_t)i * STACK_BYTES, 4096, PROT_NONE);   // guard page below each stack
    dyn_smem = (char *)aligned_alloc(1024, DYN_SMEM_BYTES);
}

static void yield(int state)
{
    Fiber &f = fibers[cur];
    f.state = state;
#if EMU_TSAN
    char *token = state == AT_BLOCK ? &tsan_block_token : &tsan_warp_token[cur >> 5];
    __tsan_release(token);                    // all lanes arrive (release) before any of them leaves (acquire)
    __tsan_switch_to_fiber(tsan_sched, __tsan_switch_to_fiber_no_sync);
#endif
#if EMU_ASAN
    __sanitizer_start_switch_fiber(&f.fake, sched_bottom, sched_size);
#endif
    emu_switch(&f.sp, sched_sp);
#if EMU_ASAN
    __sanitizer_finish_switch_fiber(f.fake, &sched_bottom, &sched_size);
#endif
#if EMU_TSAN
    __tsan_acquire(token);
#endif
}

void block_barrier() { yield(AT_BLOCK); }
void warp_barrier() { yield(AT_WARP); }
int lane_id() { return cur & 31; }
uint64_t &slot_of_lane(int lane) { return fibers[(cur & ~31) + lane].slot; }
bool lane_alive(int lane) { const int t = (cur & ~31) + lane; return t < n_threads && fibers[t].state != DONE; }

// VFSMS_EMU_ORDER=reverse: warps and lanes are scheduled last-to-first (and CTAs of a grid last-to-first).  Results that depend
// on the order in which threads run between two barriers -- a missing __syncthreads / __syncwarp, an atomic arrival order that
// leaks into the output -- differ between the two schedules.
static bool reverse_order = false;

static void run_fiber(int t)
{
    cur = t;
    threadIdx = fibers[t].tid;
#if EMU_TSAN
    __tsan_switch_to_fiber(tsan_fibers[t], __tsan_switch_to_fiber_no_sync);
#endif
#if EMU_ASAN
    void *fake = nullptr;
    __sanitizer_start_switch_fiber(&fake, stacks + (size_t)t * STACK_BYTES + 4096, STACK_BYTES - 4096);
#endif
    emu_switch(&sched_sp, fibers[t].sp);
#if EMU_ASAN
    __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#endif
    cur = -1;
}

static void run_block(const std::function<void()> &body)
{
    n_threads = (int)(blockDim.x * blockDim.y * blockDim.z);
    if (n_threads > MAX_THREADS || n_threads <= 0) { fprintf(stderr, "emu: block of %d threads\n", n_threads); abort(); }
    cur_body = &body;
#if EMU_TSAN
    if (!tsan_sched) tsan_sched = __tsan_get_current_fiber();
    for (int t = 0; t < n_threads; t++) if (!tsan_fibers[t]) tsan_fibers[t] = __tsan_create_fiber(0);
    __tsan_release(&tsan_launch_token);
#endif
    for (int t = 0; t < n_threads; t++) {
        Fiber &f = fibers[t];
        uint64_t *sp = (uint64_t *)(stacks + (size_t)(t + 1) * STACK_BYTES);
        *--sp = 0;                               // keeps rsp % 16 == 8 at the entry of fiber_main, as after a call
        *--sp = (uint64_t)(uintptr_t)&fiber_main;
        for (int r = 0; r < 6; r++) *--sp = 0;
        f.sp = sp;
        f.state = RUN;
        f.slot = 0;
        f.tid.x = t % blockDim.x;
        f.tid.y = (t / blockDim.x) % blockDim.y;
        f.tid.z = t / (blockDim.x * blockDim.y);
    }
    const int n_warps = (n_threads + 31) / 32;
    int live = n_threads;
    while (live > 0) {
        bool progressed = false;
        for (int wi = 0; wi < n_warps; wi++) {
            const int w = reverse_order ? n_warps - 1 - wi : wi;
            const int t0 = w * 32, t1 = t0 + 32 < n_threads ? t0 + 32 : n_threads;
            for (;;) {   // a warp runs until all of its lanes wait at a block barrier or have finished
                bool ran = false;
                for (int ti = t0; ti < t1; ti++) {
                    const int t = reverse_order ? t1 - 1 - (ti - t0) : ti;
                    if (fibers[t].state == RUN) { run_fiber(t); ran = true; if (fibers[t].state == DONE) live--; }
                }
                int at_warp = 0, alive = 0;
                for (int t = t0; t < t1; t++) { alive += fibers[t].state != DONE; at_warp += fibers[t].state == AT_WARP; }
                if (at_warp && at_warp == alive) { for (int t = t0; t < t1; t++) if (fibers[t].state == AT_WARP) fibers[t].state = RUN; ran = true; }
                progressed |= ran;
                if (!ran) break;
            }
        }
        int at_block = 0, alive = 0;
        for (int t = 0; t < n_threads; t++) { alive += fibers[t].state != DONE; at_block += fibers[t].state == AT_BLOCK; }
        if (alive && at_block == alive) { for (int t = 0; t < n_threads; t++) if (fibers[t].state == AT_BLOCK) fibers[t].state = RUN; progressed = true; }
        if (!progressed && live > 0) {
            fprintf(stderr, "emu: deadlock in block (%u,%u,%u): %d live threads, %d at __syncthreads, the rest at a warp collective "
                            "with divergent lanes\n", blockIdx.x, blockIdx.y, blockIdx.z, alive, at_block);
            abort();
        }
    }
#if EMU_TSAN
    __tsan_acquire(&tsan_done_token);         // the host (and, through the launch token, the next block) see everything this block did
#endif
}

static cudaError_t last_error = cudaSuccess;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body)
{
    init_once();
    const char *order = getenv("VFSMS_EMU_ORDER");
    reverse_order = order && !strcmp(order, "reverse");
    if (smem > DYN_SMEM_BYTES) { fprintf(stderr, "emu: %zu B of dynamic shared memory\n", smem); last_error = cudaErrorInvalidValue; return; }
    gridDim = grid;
    blockDim = block;
    for (unsigned z = 0; z < grid.z; z++)
        for (unsigned y = 0; y < grid.y; y++)
            for (unsigned x = 0; x < grid.x; x++) {
                blockIdx.x = reverse_order ? grid.x - 1 - x : x;
                blockIdx.y = reverse_order ? grid.y - 1 - y : y;
                blockIdx.z = reverse_order ? grid.z - 1 - z : z;
                run_block(body);
            }
}

}   // namespace emu

// ---------------------------------------------------------------- CUDA runtime: "device" memory is host memory, streams are synchronous
extern "C" {

cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { cudaError_t e = emu::last_error; emu::last_error = cudaSuccess; return e; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *v, enum cudaDeviceAttr attr, int)
{
    // VFSMS_EMU_TEX_ROWS: a small 2-D texture height limit, so that the grouping of stacked textures is exercised by small batches
    const char *tex_rows = getenv("VFSMS_EMU_TEX_ROWS");
    *v = attr == cudaDevAttrMaxTexture2DLinearHeight ? (tex_rows ? atoi(tex_rows) : 65000) : (attr == cudaDevAttrMultiProcessorCount ? 4 : 0);
    return cudaSuccess;
}
cudaError_t cudaGetDeviceProperties(struct cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof(*p));
    strcpy(p->name, "cuda_emu (CPU)");
    p->major = 10; p->minor = 0;
    p->multiProcessorCount = 4;          // keeps the grids of persistent kernels small
    p->sharedMemPerBlockOptin = 227 << 10;
    p->totalGlobalMem = (size_t)8 << 30;
    return cudaSuccess;
}
cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); if (*p) memset(*p, 0xCD, n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, enum cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, enum cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, enum cudaMemcpyKind, cudaStream_t)
{
    for (size_t r = 0; r < h; r++) memmove((char *)d + r * dp, (const char *)s + r * sp, w);
    return cudaSuccess;
}
cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, enum cudaMemcpyKind k) { return cudaMemcpy2DAsync(d, dp, s, sp, w, h, k, 0); }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemset2DAsync(void *d, size_t p, int v, size_t w, size_t h, cudaStream_t)
{
    for (size_t r = 0; r < h; r++) memset((char *)d + r * p, v, w);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)(uintptr_t)0x51; return cudaSuccess; }
cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = (cudaStream_t)(uintptr_t)0x51; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)(uintptr_t)0xE1; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)(uintptr_t)0xE1; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
cudaError_t cudaMemcpyToSymbol(const void *sym, const void *src, size_t n, size_t off, enum cudaMemcpyKind) { memcpy((char *)sym + off, src, n); return cudaSuccess; }
struct cudaChannelFormatDesc cudaCreateChannelDesc(int x, int y, int z, int w, enum cudaChannelFormatKind f)
{
    struct cudaChannelFormatDesc d;
    d.x = x; d.y = y; d.z = z; d.w = w; d.f = f;
    return d;
}
cudaError_t cudaFuncSetAttribute(const void *, enum cudaFuncAttribute, int) { return cudaSuccess; }

cudaError_t cudaCreateTextureObject(cudaTextureObject_t *obj, const struct cudaResourceDesc *res, const struct cudaTextureDesc *td,
                                    const struct cudaResourceViewDesc *)
{
    if (res->resType != cudaResourceTypePitch2D || td->normalizedCoords ||
        td->addressMode[0] != cudaAddressModeClamp || td->addressMode[1] != cudaAddressModeClamp || td->readMode != cudaReadModeElementType)
        return cudaErrorNotSupported;
    emu::Tex *t = new emu::Tex;
    t->base = (const char *)res->res.pitch2D.devPtr;
    t->width = (int)res->res.pitch2D.width;
    t->height = (int)res->res.pitch2D.height;
    t->pitch = res->res.pitch2D.pitchInBytes;
    t->elem = res->res.pitch2D.desc.x / 8;
    *obj = (cudaTextureObject_t)(uintptr_t)t;
    return cudaSuccess;
}
cudaError_t cudaDestroyTextureObject(cudaTextureObject_t obj) { delete (emu::Tex *)(uintptr_t)obj; return cudaSuccess; }

}   // extern "C"

// ---------------------------------------------------------------- cuFFT: plain O(n^2)-per-line DFTs in double (test sizes are small)
namespace {
struct Plan { int rank = 0, n[3] = {0, 0, 0}, batch = 0; cufftType type = CUFFT_D2Z; };
std::map<cufftHandle, Plan> plans;
int next_plan = 1;
typedef std::complex<double> cd;

void dft_line(const cd *in, cd *out, int n, int stride, int sign, const std::vector<cd> &tw)
{
    for (int k = 0; k < n; k++) {
        cd acc = 0;
        for (int j = 0; j < n; j++) acc += in[(size_t)j * stride] * tw[(size_t)((long long)j * k % n)];
        out[k] = acc;
    }
    (void)sign;
}
std::vector<cd> twiddles(int n, int sign)
{
    std::vector<cd> tw(n);
    for (int j = 0; j < n; j++) tw[j] = std::polar(1.0, sign * 2.0 * M_PI * j / n);
    return tw;
}
}   // namespace

extern "C" {
cufftResult cufftCreate(cufftHandle *h) { *h = next_plan++; plans[*h] = Plan(); return CUFFT_SUCCESS; }
cufftResult cufftDestroy(cufftHandle h) { plans.erase(h); return CUFFT_SUCCESS; }
cufftResult cufftSetAutoAllocation(cufftHandle, int) { return CUFFT_SUCCESS; }
cufftResult cufftSetWorkArea(cufftHandle, void *) { return CUFFT_SUCCESS; }
cufftResult cufftSetStream(cufftHandle, cudaStream_t) { return CUFFT_SUCCESS; }
cufftResult cufftMakePlanMany(cufftHandle h, int rank, int *n, int *inembed, int, int, int *onembed, int, int, cufftType type, int batch, size_t *work)
{
    if (rank != 2 || inembed || onembed || (type != CUFFT_D2Z && type != CUFFT_Z2D)) return CUFFT_NOT_SUPPORTED;
    Plan &p = plans[h];
    p.rank = rank; p.n[0] = n[0]; p.n[1] = n[1]; p.batch = batch; p.type = type;
    if (work) *work = 0;
    return CUFFT_SUCCESS;
}
// real [M][N] -> complex [M][N/2+1], forward, unnormalised
cufftResult cufftExecD2Z(cufftHandle h, cufftDoubleReal *in, cufftDoubleComplex *out)
{
    const Plan &p = plans[h];
    const int M = p.n[0], N = p.n[1], H = N / 2 + 1;
    const std::vector<cd> twN = twiddles(N, -1), twM = twiddles(M, -1);
    std::vector<cd> rowin(N), rowout(N), tmp((size_t)M * H), colout(M);
    for (int b = 0; b < p.batch; b++) {
        const double *src = in + (size_t)b * M * N;
        cd *dst = (cd *)out + (size_t)b * M * H;
        for (int r = 0; r < M; r++) {
            for (int c = 0; c < N; c++) rowin[c] = src[(size_t)r * N + c];
            dft_line(rowin.data(), rowout.data(), N, 1, -1, twN);
            for (int c = 0; c < H; c++) tmp[(size_t)r * H + c] = rowout[c];
        }
        for (int c = 0; c < H; c++) {
            dft_line(tmp.data() + c, colout.data(), M, H, -1, twM);
            for (int r = 0; r < M; r++) dst[(size_t)r * H + c] = colout[r];
        }
    }
    return CUFFT_SUCCESS;
}
// complex [M][N/2+1] (Hermitian half) -> real [M][N], inverse, unnormalised
cufftResult cufftExecZ2D(cufftHandle h, cufftDoubleComplex *in, cufftDoubleReal *out)
{
    const Plan &p = plans[h];
    const int M = p.n[0], N = p.n[1], H = N / 2 + 1;
    const std::vector<cd> twN = twiddles(N, +1), twM = twiddles(M, +1);
    std::vector<cd> tmp((size_t)M * H), colout(M), rowin(N), rowout(N);
    for (int b = 0; b < p.batch; b++) {
        const cd *src = (const cd *)in + (size_t)b * M * H;
        double *dst = out + (size_t)b * M * N;
        for (int c = 0; c < H; c++) {
            dft_line(src + c, colout.data(), M, H, +1, twM);
            for (int r = 0; r < M; r++) tmp[(size_t)r * H + c] = colout[r];
        }
        for (int r = 0; r < M; r++) {
            for (int c = 0; c < H; c++) rowin[c] = tmp[(size_t)r * H + c];
            for (int c = H; c < N; c++) rowin[c] = std::conj(tmp[(size_t)r * H + (N - c)]);
            dft_line(rowin.data(), rowout.data(), N, 1, +1, twN);
            for (int c = 0; c < N; c++) dst[(size_t)r * N + c] = rowout[c].real();
        }
    }
    return CUFFT_SUCCESS;
}
}   // extern "C"

// ---------------------------------------------------------------- the tensor-core matcher is inline PTX: not emulated
#ifndef EMU_STANDALONE
int match_tc_batch(vfsms_ctx *, const float *, const int32_t *, int, const float *, const int32_t *, int, int, int, int, int32_t *, float *, cudaStream_t)
{
    vfsms_set_error("cuda_emu: match_tc.cu (tcgen05 / TMA inline PTX) is not emulated; select the exact SIMT matcher with vfsms_set_matcher(ctx, 1)");
    return VFSMS_E_UNSUPPORTED;
}
#endif
