// emu.h -- TEST INFRASTRUCTURE ONLY: a minimal CUDA-on-CPU execution model, so that the kernels of imagestitch_b200/csrc/*.cu
// can be exercised by the CPU test-suite (tests/test_kernels_emulated_cpu.py) without a GPU.  Nothing in the product links this.
//
// * The .cu sources are compiled by g++ after a textual rewrite of the <<<...>>> launches (tests/cuda_emu/build_emu.py);
//   this header is force-included and supplies the device-side vocabulary (threadIdx, __shared__, warp collectives,
//   atomics, conversion intrinsics, tex2Dgather) on top of the toolkit's own host headers (types and API prototypes).
// * A kernel launch runs block after block; every CUDA thread of a block is a ucontext fiber.  Fibers switch only at
//   __syncthreads() and at warp collectives (__shfl*_sync, __ballot_sync, __syncwarp), which are implemented as barriers over
//   the warp's live lanes + a per-lane exchange slot: lock-step semantics without OS threads, deterministic.
// * emu.cpp implements the handful of CUDA runtime calls the host code makes (cudaMalloc = calloc, copies = memcpy,
//   texture objects = a small descriptor, streams / events = no-ops) and stubs cuFFT (phase correlation is not emulated).
// It checks kernel LOGIC (indexing, reductions, work distribution, exact arithmetic order); it says nothing about timing,
// memory-model races or the tensor-core path (match_tc.cu is stubbed: inline PTX).
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>

#undef __global__
#define __global__
#undef __device__
#define __device__
#undef __host__
#define __host__
#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __grid_constant__
#define __grid_constant__
#undef __constant__
#define __constant__

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
static const int warpSize = 32;

namespace emu {
extern char *dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
void block_barrier();
void warp_barrier();
int lane_id();
uint64_t &slot_of_lane(int lane);       // exchange slot of a lane of the current warp
bool lane_alive(int lane);
struct Tex { const char *base; int width, height; size_t pitch; int elem; };
}

// ---------------------------------------------------------------- synchronisation / warp collectives
static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }

template <class T> static inline T emu_exchange(T v, int src)
{
    static_assert(sizeof(T) <= 8, "shuffle payload");
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    emu::slot_of_lane(emu::lane_id()) = bits;
    emu::warp_barrier();
    T r = v;
    if (src >= 0 && src < 32 && emu::lane_alive(src)) { const uint64_t o = emu::slot_of_lane(src); memcpy(&r, &o, sizeof(T)); }
    emu::warp_barrier();
    return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu_exchange(v, src & 31); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu_exchange(v, emu::lane_id() ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { const int s = emu::lane_id() + (int)d; return emu_exchange(v, s < 32 ? s : emu::lane_id()); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { const int s = emu::lane_id() - (int)d; return emu_exchange(v, s >= 0 ? s : emu::lane_id()); }
static inline unsigned __ballot_sync(unsigned, int pred)
{
    emu::slot_of_lane(emu::lane_id()) = pred ? 1 : 0;
    emu::warp_barrier();
    unsigned r = 0;
    for (int l = 0; l < 32; l++) if (emu::lane_alive(l) && emu::slot_of_lane(l)) r |= 1u << l;
    emu::warp_barrier();
    return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, !p) == 0; }

// ---------------------------------------------------------------- atomics (one OS thread: plain read-modify-write)
// (relaxed __atomic builtins: the same plain read-modify-write for one OS thread, but ThreadSanitizer knows they are atomics)
template <class T> struct emu_atomic_int { typedef T type; };
template <> struct emu_atomic_int<float> { typedef unsigned type; };
template <> struct emu_atomic_int<double> { typedef unsigned long long type; };
template <class T> static inline T emu_aload(T *p) { typename emu_atomic_int<T>::type b = __atomic_load_n((typename emu_atomic_int<T>::type *)p, __ATOMIC_RELAXED); T v; memcpy(&v, &b, sizeof(T)); return v; }
template <class T> static inline void emu_astore(T *p, T v) { typename emu_atomic_int<T>::type b; memcpy(&b, &v, sizeof(T)); __atomic_store_n((typename emu_atomic_int<T>::type *)p, b, __ATOMIC_RELAXED); }
template <class T, class U> static inline T atomicAdd(T *p, U v) { T o = emu_aload(p); emu_astore(p, (T)(o + (T)v)); return o; }
template <class T, class U> static inline T atomicMin(T *p, U v) { T o = emu_aload(p); if ((T)v < o) emu_astore(p, (T)v); return o; }
template <class T, class U> static inline T atomicMax(T *p, U v) { T o = emu_aload(p); if ((T)v > o) emu_astore(p, (T)v); return o; }
template <class T, class U> static inline T atomicOr(T *p, U v) { T o = emu_aload(p); emu_astore(p, (T)(o | (T)v)); return o; }
template <class T, class U> static inline T atomicExch(T *p, U v) { T o = emu_aload(p); emu_astore(p, (T)v); return o; }
template <class T, class U, class V> static inline T atomicCAS(T *p, U cmp, V v) { T o = emu_aload(p); if (o == (T)cmp) emu_astore(p, (T)v); return o; }
template <class T> static inline unsigned atomicInc(T *p, unsigned lim) { unsigned o = emu_aload(p); emu_astore(p, (T)(o >= lim ? 0 : o + 1)); return o; }

// ---------------------------------------------------------------- intrinsics
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int __float2int_rn(float v) { return (int)lrintf(v); }
static inline float __int2float_rn(int v) { return (float)v; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __ull2float_rn(unsigned long long v) { return (float)v; }
static inline float __uint2float_rn(unsigned v) { return (float)v; }      // round to nearest even (the default FP environment)
static inline int __float2int_rd(float v) { return (int)floorf(v); }
static inline int __float2int_rz(float v) { return (int)v; }
static inline int __double2int_rn(double v) { return (int)lrint(v); }
static inline int __double2int_rd(double v) { return (int)floor(v); }
static inline int __double2int_ru(double v) { return (int)ceil(v); }
static inline int __double2loint(double v) { uint64_t b; memcpy(&b, &v, 8); return (int)(uint32_t)b; }
static inline int __double2hiint(double v) { uint64_t b; memcpy(&b, &v, 8); return (int)(uint32_t)(b >> 32); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
    const uint64_t all = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(all >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0;
        r |= byte << (8 * i);
    }
    return r;
}
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float v) { return 1.0f / sqrtf(v); }
static inline float __saturatef(float v) { return v < 0 ? 0 : (v > 1 ? 1 : v); }
static inline void __threadfence() {}
static inline void __trap() { fprintf(stderr, "emu: __trap()\n"); abort(); }

// CUDA's integer / float min / max overload set (device code calls them unqualified)
#define EMU_MINMAX(T) static inline T min(T a, T b) { return b < a ? b : a; } static inline T max(T a, T b) { return a < b ? b : a; }
EMU_MINMAX(int) EMU_MINMAX(unsigned) EMU_MINMAX(long) EMU_MINMAX(unsigned long) EMU_MINMAX(long long) EMU_MINMAX(unsigned long long)
EMU_MINMAX(float) EMU_MINMAX(double)
#undef EMU_MINMAX
static inline long long min(long long a, int b) { return a < b ? a : b; }
static inline long long min(int a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, int b) { return a > b ? a : b; }
static inline long long max(int a, long long b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
static inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
static inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
static inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
static inline float min(float a, int b) { return fminf(a, (float)b); }
static inline float max(float a, int b) { return fmaxf(a, (float)b); }
static inline double min(double a, float b) { return fmin(a, (double)b); }
static inline double max(double a, float b) { return fmax(a, (double)b); }
static inline double min(float a, double b) { return fmin((double)a, b); }
static inline double max(float a, double b) { return fmax((double)a, b); }

// ---------------------------------------------------------------- textures: point-sampled pitch-2D, clamp addressing
template <class T> static inline T tex2Dgather(cudaTextureObject_t obj, float x, float y, int comp = 0);
template <class E> static inline float emu_texel(const emu::Tex *t, int i, int j)
{
    i = i < 0 ? 0 : (i >= t->width ? t->width - 1 : i);
    j = j < 0 ? 0 : (j >= t->height ? t->height - 1 : j);
    return (float)*(const E *)(t->base + (size_t)j * t->pitch + (size_t)i * sizeof(E));
}
static inline float emu_texel_any(const emu::Tex *t, int i, int j)
{
    return t->elem == 4 ? emu_texel<float>(t, i, j) : emu_texel<unsigned char>(t, i, j);
}
// the four texels of the bilinear footprint of (x, y): .x = (i0, j0+1), .y = (i0+1, j0+1), .z = (i0+1, j0), .w = (i0, j0)
template <> inline float4 tex2Dgather<float4>(cudaTextureObject_t obj, float x, float y, int)
{
    const emu::Tex *t = (const emu::Tex *)(uintptr_t)obj;
    const int i0 = (int)floorf(x - 0.5f), j0 = (int)floorf(y - 0.5f);
    return make_float4(emu_texel_any(t, i0, j0 + 1), emu_texel_any(t, i0 + 1, j0 + 1), emu_texel_any(t, i0 + 1, j0), emu_texel_any(t, i0, j0));
}
// bilinear filtering as the texture unit does it: 8 fractional bits for the weights (CUDA programming guide, "linear filtering")
template <class T> static inline T tex2D(cudaTextureObject_t obj, float x, float y);
template <> inline float tex2D<float>(cudaTextureObject_t obj, float x, float y)
{
    const emu::Tex *t = (const emu::Tex *)(uintptr_t)obj;
    const float xb = x - 0.5f, yb = y - 0.5f;
    const int i0 = (int)floorf(xb), j0 = (int)floorf(yb);
    const float a = floorf((xb - (float)i0) * 256.0f + 0.5f) / 256.0f, b = floorf((yb - (float)j0) * 256.0f + 0.5f) / 256.0f;
    return (1 - a) * (1 - b) * emu_texel_any(t, i0, j0) + a * (1 - b) * emu_texel_any(t, i0 + 1, j0) +
           (1 - a) * b * emu_texel_any(t, i0, j0 + 1) + a * b * emu_texel_any(t, i0 + 1, j0 + 1);
}
template <> inline uchar4 tex2Dgather<uchar4>(cudaTextureObject_t obj, float x, float y, int)
{
    const float4 g = tex2Dgather<float4>(obj, x, y, 0);
    return make_uchar4((unsigned char)g.x, (unsigned char)g.y, (unsigned char)g.z, (unsigned char)g.w);
}

template <class T> static inline cudaError_t cudaFuncSetAttribute(T *, enum cudaFuncAttribute, int) { return cudaSuccess; }
// `kernel<<<grid, block, smem, stream>>>(args)` after build_emu.translate
namespace emu {
static inline void launch_cfg(const std::function<void()> &body, dim3 grid, dim3 block, size_t smem = 0, cudaStream_t = 0) { launch(grid, block, smem, body); }
}
