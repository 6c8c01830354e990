// match_tc.cu -- tensor-core brute-force matcher (tcgen05 / TMEM / TMA), exact results.  sm_100a only.
//
// Replaces the O(nA*nB*D) part of Method.matchDescriptors (ImageUtility.py:278-309, BF L2 kNN(2)).
//
//   1. prep_split_kernel      fp32 descriptors -> 16-bit operand rows whose product is  ||b||^2 - 2 a.b  (nb = ||b||^2):
//                               fp16 x1 (default)  query [ a | 1, 1, 0 ... ]               train [ -2b | nb_hi, nb_lo, 0 ... ]
//                               fp16 x2            query [ a_hi | a_lo | 1, 1 ... ]        train [ -2b | -2b | nb ... ]
//                               bf16 x3            query [ a_hi | a_hi | a_lo | 1, 1 ... ] train [ -2b_hi | -2b_lo | -2b_hi | nb ... ]
//                             K' = terms * D + 16 columns take part (only the used 16-column steps of the last k-block are
//                             issued).  The kernel also measures the rounding error norms of the rows it writes.
//   2. match_tc_pair_kernel   persistent warp-specialised GEMM on CTA pairs (the default): TMA (SWIZZLE_128B) -> smem ->
//                             tcgen05.mma.cta_group::2 (256 x 256 x 16 per step, kind::f16, fp32 accumulate in TMEM, two
//                             accumulator buffers) -> epilogue warps read TMEM with tcgen05.ld and keep a running top-4 of
//                             (score | column) keys per query row and column half; the merge network runs only for groups
//                             of four scores in which some row of the warp can still improve.  Each CTA's 128 x K' query
//                             tile stays resident in shared memory while its half of every train tile streams through a
//                             7-stage mbarrier ring.  Nothing but 8 candidates per query is written to HBM.
//      match_tc_kernel        the same on single CTAs (128 x 128 tiles, cta_group::1); kept for verification and as the
//                             subject of the -DTC_TIMING probes that located the shared-memory operand bottleneck.
//   3. rescore_kernel         exact fp32 distances (serial k order, no FMA: the CPU value) of the candidates that can be
//                             one of the best two, best two by (distance, index), plus a guard: if the second best exact
//                             score is not separated from the worst kept key by more than the operand error bound, the
//                             query is flagged ...
//   4. fallback_scan_kernel   ... and rescanned exactly against every train row, all flagged queries of a pair together.
//                             Results are therefore identical to the exact SIMT kernel (match.cu) by construction, not by luck.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>

#define TC_M 128
#define TC_N 128
#define TC_KB 64                 // bf16 elements per k-block = 128 bytes = one swizzle row
#define TC_STAGES 7
#define TC_TOPK 4
#ifndef TC_EPI_GROUPS
#define TC_EPI_GROUPS 2          // epilogue warpgroups; group g scans its 1 / TC_EPI_GROUPS share of the columns of every tile.
#endif                           // 4 (16 epilogue warps, 16 candidates per query) measured slower: GEMM 1.07 vs 0.95 ms, rescoring 0.22 vs 0.13 ms
#define TC_MAX_KBLOCKS 7         // K' <= 448  (D <= 128)
#define TC_THREADS (128 + 128 * TC_EPI_GROUPS)
// Warp roles.  The schedulers favour the HIGHEST warp id of a sub-partition, so the two latency-critical single-thread
// warps (TMA producer, MMA issuer) sit above the ALU-heavy epilogue warps they share a scheduler with: as warps 0/1 the
// issuer was starved and every UMMA took ~156 cycles instead of 64 (measured with clock64 around its waits).
#define TC_EPI_WARPS (4 * TC_EPI_GROUPS)          // warps 0 .. 7: epilogue (TMEM lane quarter = warp & 3, column half = warp >> 2)
#define TC_WARP_TMA (TC_EPI_WARPS + 0)
#define TC_WARP_MMA (TC_EPI_WARPS + 1)
#define TC_WARP_ALLOC (TC_EPI_WARPS + 2)

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)     // arrives on `bar` when all prior MMAs of this thread finished
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, version 1):
//   start address >> 4 | LBO (ignored for swizzled K-major; 1) | SBO = 8 rows * 128 B = 1024 B | layout type 2 (128B swizzle)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor: D = f32, both operands K-major, N = 128, M = 128; `fmt` = TC_FMT_BF16 or TC_FMT_F16 (A and B format fields)
#define TC_FMT_BF16 ((1u << 7) | (1u << 10))
#define TC_FMT_F16 0u
__device__ __forceinline__ uint32_t make_idesc(uint32_t fmt)
{
    return (1u << 4) | fmt | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

// ---------------------------------------------------------------- 1. split operand rows
// grid (ceil(cap/8), images), 256 threads: one warp per row.  side 0 = query layout, 1 = train layout.
// terms 3: split-bf16 rows (header);  terms 2 / 1: fp16 rows
//     query  [ a_hi | a_lo | 1, 1, 0 ... ]      resp.  [ a_hi | 1, 1, 0 ... ]
//     train  [ B    | B    | nb_hi, nb_lo ... ] resp.  [ B    | nb_hi, nb_lo ... ]        B = fp16(-2 b)
// `used` = terms * dim + 16 columns take part in the product; the row is padded to kprime (a multiple of 64) with zeros.
// fp16 has no room for large values: a row whose squared norm exceeds TC_F16_NORM_MAX (or is not finite) raises the pair's
// overflow flag and rescore_kernel hands every query of that pair to the exact rescan.
#define TC_F16_NORM_MAX 1e4f
#define TC_F16_PAD_NORM 60000.f      // fp16-representable, above every score a pair without the overflow flag can produce
__global__ void __launch_bounds__(256) prep_split_kernel(const float *__restrict__ desc, const int32_t *n_ptr, int n_stride,
                                                         int cap, int dim, int kprime, int terms, int side, int img0,
                                                         uint16_t *out, float *norms, float *errs, int32_t *ovf)
{
    const int b = blockIdx.y;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= cap) return;
    const int n = n_ptr[(size_t)(img0 + b) * n_stride];
    const float *src = desc + ((size_t)(img0 + b) * cap + row) * dim;
    uint16_t *dst = out + ((size_t)b * cap + row) * kprime;
    const bool valid = row < n;
    const bool f16 = terms < 3;
    float nrm = 0.f, er = 0.f;      // er: squared norm of what the stored operand row lost (fp16 schemes; both differences are exact)
    for (int k = lane; k < dim; k += 32) {
        const float v = valid ? src[k] : 0.f;
        nrm += v * v;
        if (f16) {
            if (side == 0) {
                const __half hi = __float2half_rn(v);
                float e = v - __half2float(hi);
                dst[k] = __half_as_ushort(hi);
                if (terms == 2) {
                    const __half lo = __float2half_rn(e);
                    dst[dim + k] = __half_as_ushort(lo);
                    e -= __half2float(lo);
                }
                er += e * e;
            } else {
                const __half m = __float2half_rn(-2.f * v);
                const float e = __half2float(m) + 2.f * v;
                dst[k] = __half_as_ushort(m);
                if (terms == 2) dst[dim + k] = __half_as_ushort(m);
                er += e * e;
            }
        } else {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
            if (side == 0) { dst[k] = __bfloat16_as_ushort(hi); dst[dim + k] = __bfloat16_as_ushort(hi); dst[2 * dim + k] = __bfloat16_as_ushort(lo); }
            else {
                const uint16_t mhi = __bfloat16_as_ushort(__float2bfloat16_rn(-2.f * __bfloat162float(hi)));
                const uint16_t mlo = __bfloat16_as_ushort(__float2bfloat16_rn(-2.f * __bfloat162float(lo)));
                dst[k] = mhi; dst[dim + k] = mlo; dst[2 * dim + k] = mhi;
            }
        }
    }
    for (int o = 16; o; o >>= 1) { nrm += __shfl_xor_sync(0xffffffffu, nrm, o); er += __shfl_xor_sync(0xffffffffu, er, o); }
    if (f16 && valid && lane == 0 && !(nrm <= TC_F16_NORM_MAX)) ovf[b] = 1;
    for (int k = terms * dim + lane; k < kprime; k += 32) {
        float v = 0.f;
        const int e = k - terms * dim;
        if (side == 0) v = (e < 2) ? 1.f : 0.f;
        else if (f16) {
            const float nb = valid ? fminf(nrm, TC_F16_NORM_MAX) : TC_F16_PAD_NORM;       // padding train rows can never be selected
            const float hi = __half2float(__float2half_rn(nb));
            v = e == 0 ? hi : (e == 1 ? (valid ? nb - hi : 0.f) : 0.f);
        } else {
            const float nb = valid ? nrm : 1e30f;
            const float hi = __bfloat162float(__float2bfloat16_rn(nb));
            v = e == 0 ? hi : (e == 1 ? (valid ? nb - hi : 0.f) : 0.f);
        }
        dst[k] = f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
    }
    if (lane == 0) { norms[(size_t)b * cap + row] = valid ? nrm : 0.f; errs[(size_t)b * cap + row] = valid ? er : 0.f; }
}

// ---------------------------------------------------------------- 2. the GEMM + running top-K
struct TcShared {
    uint64_t a_full, a_empty;
    uint64_t b_full[TC_STAGES], b_empty[TC_STAGES];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

// ---- epilogue: branch-free top-4 on KEYS.
// A key is the fp32 score with its low 7 mantissa bits replaced by the column inside the warp's share of the tile, so the
// selection network is pure FMNMX (no index registers, no divergence): ordering by key differs from ordering by score by
// at most 2^-16 |score|, which the rescoring guard adds to its error bound.  Four new keys at a time: sort them (5
// comparators), take the lower half of the bitonic sequence (list ascending, new keys descending), re-sort (4 comparators)
// = 5.5 FMNMX + 1 LOP3 per score.  The old data-dependent insert cost ~46 issue slots per score and bounded the GEMM.
#define TC_KEY_BITS 7
#define TC_KEY_MASK (~((1u << TC_KEY_BITS) - 1u))
__device__ __forceinline__ void cmpx(float &a, float &b) { const float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi; }
__device__ __forceinline__ void top4_merge4(float (&m)[4], float a, float b, float c, float d)
{
    cmpx(a, b); cmpx(c, d); cmpx(a, c); cmpx(b, d); cmpx(b, c);                       // a <= b <= c <= d
    float c0 = fminf(m[0], d), c1 = fminf(m[1], c), c2 = fminf(m[2], b), c3 = fminf(m[3], a);
    cmpx(c0, c2); cmpx(c1, c3); cmpx(c0, c1); cmpx(c2, c3);
    m[0] = c0; m[1] = c1; m[2] = c2; m[3] = c3;
}
__device__ __forceinline__ float fmin3(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));      // FMNMX3
    return r;
}
// Four scores at a time.  Once a row has seen a few hundred train columns almost no score can still enter its list
// (chance ~4/n per column), so the merge network runs only when some lane of the warp has a score below `thr`, the
// smaller of the tile list's and the item list's fourth key: 2 FMNMX(3) + FSETP + VOTE + BRA per skipped group instead
// of 22 FMNMX + 4 LOP3.  A skipped score is >= thr >= the final fourth key, which is all the rescoring guard relies on.
// The eight votes of a 32-column block are taken together against the threshold at its start (a stale threshold only
// lets a few more groups through): one min -> compare -> vote latency per block instead of one per group.
__device__ __forceinline__ void tile_top4(const uint32_t (&v)[32], int col_base, float (&l)[4], float &thr)
{
    bool hit[8];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float m = fminf(fmin3(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2])), __uint_as_float(v[j + 3]));
        hit[j >> 2] = __any_sync(0xffffffffu, m < thr);
    }
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        if (hit[j >> 2]) {
            top4_merge4(l, __uint_as_float((v[j] & TC_KEY_MASK) | (uint32_t)(col_base + j)), __uint_as_float((v[j + 1] & TC_KEY_MASK) | (uint32_t)(col_base + j + 1)),
                        __uint_as_float((v[j + 2] & TC_KEY_MASK) | (uint32_t)(col_base + j + 2)), __uint_as_float((v[j + 3] & TC_KEY_MASK) | (uint32_t)(col_base + j + 3)));
            thr = fminf(thr, l[3]);
        }
    }
}
// the item-wide list carries the train tile of every key beside it; a tile's four keys enter here (ascending, so the
// loop stops at the first that no lane can use)
__device__ __forceinline__ void running_insert(float s, int tile, float (&r)[4], int (&t)[4])
{
    const bool p0 = s < r[0], p1 = s < r[1], p2 = s < r[2], p3 = s < r[3];
    t[3] = p2 ? t[2] : (p3 ? tile : t[3]);
    t[2] = p1 ? t[1] : (p2 ? tile : t[2]);
    t[1] = p0 ? t[0] : (p1 ? tile : t[1]);
    t[0] = p0 ? tile : t[0];
    r[3] = fminf(r[3], fmaxf(r[2], s));
    r[2] = fminf(r[2], fmaxf(r[1], s));
    r[1] = fminf(r[1], fmaxf(r[0], s));
    r[0] = fminf(r[0], s);
}
// one accumulator tile: this warp's 32 rows x 64 columns.  The TMEM buffer is handed back to the MMA warp as soon as
// the scores are in registers.
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, uint64_t *acc_empty, int lane, int tile, float (&r)[4], int (&t)[4])
{
    static_assert(TC_TOPK == 4 && (TC_N / TC_EPI_GROUPS == 64 || TC_N / TC_EPI_GROUPS == 32), "epilogue is written for top-4 over 64- or 32-column shares");
    uint32_t v0[32], v1[32];
    tmem_ld32_issue(taddr, v0);
    if (TC_N / TC_EPI_GROUPS == 64) tmem_ld32_issue(taddr + 32, v1);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(acc_empty);
    float l[4] = { FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX };
    float thr = r[3];
    tile_top4(v0, 0, l, thr);
    if (TC_N / TC_EPI_GROUPS == 64) tile_top4(v1, 32, l, thr);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (!__any_sync(0xffffffffu, l[k] < r[3])) break;
        running_insert(l[k], tile, r, t);
    }
}
__device__ __forceinline__ void epilogue_store(const float (&r)[4], const int (&t)[4], int tile_n, int grp, size_t o, float *cand_score, int32_t *cand_idx)
{
#pragma unroll
    for (int i = 0; i < 4; i++) {
        cand_score[o + i] = r[i];
        cand_idx[o + i] = r[i] > 1e29f ? -1 : t[i] * tile_n + grp * (tile_n / TC_EPI_GROUPS) + (int)(__float_as_uint(r[i]) & ~TC_KEY_MASK);
    }
}

#ifdef TC_TIMING
__device__ unsigned long long g_tc_t[8];
#define TCT_DECL unsigned long long tw_a = 0, tw_acc = 0, tw_b = 0, tt0 = clock64(), tq
#define TCT(x, stmt) do { tq = clock64(); stmt; x += clock64() - tq; } while (0)
#else
#define TCT_DECL
#define TCT(x, stmt) stmt
#endif
__global__ void __launch_bounds__(TC_THREADS, 1) match_tc_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                 const __grid_constant__ CUtensorMap tmap_b,
                                                                 const int32_t *__restrict__ n_a_ptr, int n_a_stride,
                                                                 const int32_t *__restrict__ n_b_ptr, int n_b_stride,
                                                                 int n_pairs, int cap, int kblocks, int last_steps, uint32_t fmt, int n_splits,
                                                                 float *cand_score, int32_t *cand_idx)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);       // SWIZZLE_128B needs 1024-B alignment
    uint8_t *sA = smem;                                                 // kblocks x [128 rows x 128 B]
    uint8_t *sB = smem + (size_t)TC_MAX_KBLOCKS * TC_M * 128;           // TC_STAGES x [128 rows x 128 B]
    TcShared *sh = (TcShared *)(sB + (size_t)TC_STAGES * TC_N * 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = cap / TC_M;
    const int items = n_pairs * m_tiles * n_splits;

    if (warp == TC_WARP_MMA && lane == 0) {
        mbar_init(&sh->a_full, 1); mbar_init(&sh->a_empty, 1);
        for (int i = 0; i < TC_STAGES; i++) { mbar_init(&sh->b_full[i], 1); mbar_init(&sh->b_empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&sh->acc_full[i], 1); mbar_init(&sh->acc_empty[i], 4 * TC_EPI_GROUPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_WARP_ALLOC) tmem_alloc(&sh->tmem_base, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sh->tmem_base;

    // per-item geometry shared by all roles
    auto item_geom = [&](int item, int &p, int &mt, int &nt0, int &nt1) -> bool {
        const int s = item % n_splits;
        const int r = item / n_splits;
        mt = r % m_tiles; p = r / m_tiles;
        const int nA = n_a_ptr[(size_t)p * n_a_stride], nB = n_b_ptr[(size_t)p * n_b_stride];
        if (mt * TC_M >= nA) return false;
        const int n_tiles = (nB + TC_N - 1) / TC_N;
        const int per = (n_tiles + n_splits - 1) / n_splits;
        nt0 = s * per; nt1 = min(n_tiles, nt0 + per);
        return true;                       // nt0 >= nt1 is allowed: the epilogue still writes an empty candidate list
    };

    if (warp == TC_WARP_TMA) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, sphase = 0, a_phase = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                int p, mt, nt0, nt1;
                if (!item_geom(item, p, mt, nt0, nt1) || nt0 >= nt1) continue;
                mbar_wait(&sh->a_empty, a_phase ^ 1); a_phase ^= 1;
                mbar_expect_tx(&sh->a_full, (uint32_t)kblocks * TC_M * 128);
                for (int kb = 0; kb < kblocks; kb++)
                    tma_load_2d(sA + (size_t)kb * TC_M * 128, &tmap_a, &sh->a_full, kb * TC_KB, p * cap + mt * TC_M);
                for (int nt = nt0; nt < nt1; nt++)
                    for (int kb = 0; kb < kblocks; kb++) {
                        mbar_wait(&sh->b_empty[stage], sphase ^ 1);
                        mbar_expect_tx(&sh->b_full[stage], TC_N * 128);
                        tma_load_2d(sB + (size_t)stage * TC_N * 128, &tmap_b, &sh->b_full[stage], kb * TC_KB, p * cap + nt * TC_N);
                        if (++stage == TC_STAGES) { stage = 0; sphase ^= 1; }
                    }
            }
        }
    } else if (warp == TC_WARP_MMA) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(fmt);
            const uint64_t a_desc0 = make_sw128_desc(smem_u32(sA)), b_desc0 = make_sw128_desc(smem_u32(sB));
            uint32_t stage = 0, sphase = 0, a_phase = 0, acc = 0, acc_phase = 0;
            TCT_DECL;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                int p, mt, nt0, nt1;
                if (!item_geom(item, p, mt, nt0, nt1) || nt0 >= nt1) continue;
                TCT(tw_a, mbar_wait(&sh->a_full, a_phase)); a_phase ^= 1;
                tc_fence_after();
                for (int nt = nt0; nt < nt1; nt++) {
                    TCT(tw_acc, mbar_wait(&sh->acc_empty[acc], acc_phase ^ 1));
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * TC_N;
                    for (int kb = 0; kb < kblocks; kb++) {
                        TCT(tw_b, mbar_wait(&sh->b_full[stage], sphase));
                        tc_fence_after();
                        const uint64_t a_desc = a_desc0 + (uint64_t)(kb * (TC_M * 128 >> 4));
                        const uint64_t b_desc = b_desc0 + (uint64_t)(stage * (TC_N * 128 >> 4));
                        const int steps = kb == kblocks - 1 ? last_steps : TC_KB / 16;       // the last k-block may be partly used
#pragma unroll
                        for (int k = 0; k < TC_KB / 16; k++)      // +32 B per 16-element k step = +2 in the (addr >> 4) field
                            if (k < steps) umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) ? 1u : 0u);
                        umma_commit(&sh->b_empty[stage]);            // frees the smem stage when these MMAs are done
                        if (++stage == TC_STAGES) { stage = 0; sphase ^= 1; }
                    }
                    umma_commit(&sh->acc_full[acc]);                 // accumulator ready for the epilogue
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                umma_commit(&sh->a_empty);                           // query tile may be overwritten
            }
#ifdef TC_TIMING
            atomicAdd(&g_tc_t[0], tw_a); atomicAdd(&g_tc_t[1], tw_acc); atomicAdd(&g_tc_t[2], tw_b); atomicAdd(&g_tc_t[3], clock64() - tt0);
#endif
        }
    } else if (warp < TC_EPI_WARPS) {
        // ===================== epilogue warpgroups: TMEM -> registers -> running top-K =====================
        // Two warps per TMEM lane quarter (one per column half) so every scheduler has two epilogue warps to interleave:
        // the insert path is a dependent chain and a lone warp per scheduler cannot hide its latency.
        const int q4 = warp & 3;                                     // TMEM lane quarter this warp may access
        const int grp = warp >> 2;                             // column group
        const int row = q4 * 32 + lane;
        constexpr int GCOLS = TC_N / TC_EPI_GROUPS;
        uint32_t acc = 0, acc_phase = 0;
        TCT_DECL;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            int p, mt, nt0, nt1;
            if (!item_geom(item, p, mt, nt0, nt1)) continue;
            float r[4] = { FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX }; int t[4] = { 0, 0, 0, 0 };
            for (int nt = nt0; nt < nt1; nt++) {
                TCT(tw_acc, mbar_wait(&sh->acc_full[acc], acc_phase));
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * TC_N + grp * GCOLS;
                epilogue_tile(taddr, &sh->acc_empty[acc], lane, nt, r, t);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            const int q = mt * TC_M + row;
            const int s_id = item % n_splits;
            const size_t o = ((((size_t)p * cap + q) * n_splits + s_id) * TC_EPI_GROUPS + grp) * TC_TOPK;
            epilogue_store(r, t, TC_N, grp, o, cand_score, cand_idx);
        }
#ifdef TC_TIMING
        if (warp == 5 && lane == 0) { atomicAdd(&g_tc_t[4], tw_acc); atomicAdd(&g_tc_t[5], clock64() - tt0); }
        (void)tw_a; (void)tw_b;
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_WARP_ALLOC) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// ---------------------------------------------------------------- 2b. the same GEMM on CTA pairs (cta_group::2)
// A 1-SM UMMA of 128 x 128 x 16 reads 8 KB of operands from shared memory in its 64 issue cycles -- the whole shared
// memory bandwidth of the SM -- while TMA writes the next train block into the same memory: measured, every UMMA of
// match_tc_kernel takes ~145 cycles (tensor pipe 42 % active) no matter how deep the ring or how cheap the epilogue.
// Here two CTAs of a cluster (one TPC) run ONE 256 x 256 x 16 UMMA per step: each CTA holds its own 128 query rows and
// only HALF of the train tile (128 of 256 rows), so per SM and per 128 tensor cycles 8 KB are read and 4 KB written.
// The leader CTA's thread issues the MMAs for both; TMA transactions of both CTAs complete on the leader's barriers;
// tcgen05.commit multicasts "stage free" / "accumulator full" to both; both CTAs' epilogue warps release the
// accumulator on the leader's barrier.
#define TC2_N 256
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
// Hands an accumulator buffer back to the leader's MMA thread.  No generic-memory data travels with this signal (the
// TMEM reads are ordered by tcgen05.wait::ld + tcgen05.fence), so the default CTA-scope release is enough; the
// .release.cluster form put a cluster-scope MEMBAR (~900 cycles, 20 % of the epilogue warps' time in ncu) in front of it.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar)      // arrives on `bar` (same offset) in both CTAs
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)0x3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
match_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const int32_t *__restrict__ n_a_ptr, int n_a_stride, const int32_t *__restrict__ n_b_ptr, int n_b_stride,
                     int n_pairs, int cap, int kblocks, int last_steps, uint32_t fmt, int n_splits, float *cand_score, int32_t *cand_idx)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                                                 // this CTA's 128 query rows, all k-blocks
    uint8_t *sB = smem + (size_t)TC_MAX_KBLOCKS * TC_M * 128;           // TC_STAGES x [this CTA's 128 of the 256 train rows x 128 B]
    TcShared *sh = (TcShared *)(sB + (size_t)TC_STAGES * (TC2_N / 2) * 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cr = cluster_ctarank();
    const bool leader = cr == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int m_tiles = cap / TC_M, m_pairs = (m_tiles + 1) / 2;
    const int items = n_pairs * m_pairs * n_splits;

    if (warp == TC_WARP_MMA && lane == 0) {
        mbar_init(&sh->a_full, 1); mbar_init(&sh->a_empty, 1);
        for (int i = 0; i < TC_STAGES; i++) { mbar_init(&sh->b_full[i], 1); mbar_init(&sh->b_empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&sh->acc_full[i], 1); mbar_init(&sh->acc_empty[i], 2 * TC_EPI_WARPS); }   // both CTAs' epilogues
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_WARP_ALLOC) tmem_alloc_pair(&sh->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                       // the leader's barriers exist before the peer signals them
    tc_fence_after();
    const uint32_t tmem_base = sh->tmem_base;

    // identical control flow in both CTAs: validity is decided on the cluster's first query tile
    auto item_geom = [&](int item, int &p, int &mt, int &nt0, int &nt1) -> bool {
        const int s = item % n_splits;
        const int r = item / n_splits;
        const int mp = r % m_pairs; p = r / m_pairs;
        mt = min(mp * 2 + (int)cr, m_tiles - 1);            // an odd tail tile is computed (and stored, identically) twice
        const int nA = n_a_ptr[(size_t)p * n_a_stride], nB = n_b_ptr[(size_t)p * n_b_stride];
        if (mp * 2 * TC_M >= nA) return false;
        const int n_tiles = (nB + TC2_N - 1) / TC2_N;
        const int per = (n_tiles + n_splits - 1) / n_splits;
        nt0 = s * per; nt1 = min(n_tiles, nt0 + per);
        return true;
    };

    if (warp == TC_WARP_TMA) {
        // ===================== TMA producer (one thread per CTA): own query tile, own half of every train tile =====================
        if (lane == 0) {
            const uint32_t a_full_leader = mapa_u32(smem_u32(&sh->a_full), 0);
            uint32_t stage = 0, sphase = 0, a_phase = 0;
            for (int item = cluster_id; item < items; item += n_clusters) {
                int p, mt, nt0, nt1;
                if (!item_geom(item, p, mt, nt0, nt1) || nt0 >= nt1) continue;
                mbar_wait(&sh->a_empty, a_phase ^ 1); a_phase ^= 1;
                if (leader) mbar_expect_tx(&sh->a_full, 2u * (uint32_t)kblocks * TC_M * 128);
                for (int kb = 0; kb < kblocks; kb++)
                    tma_load_2d_pair(sA + (size_t)kb * TC_M * 128, &tmap_a, a_full_leader, kb * TC_KB, p * cap + mt * TC_M);
                for (int nt = nt0; nt < nt1; nt++)
                    for (int kb = 0; kb < kblocks; kb++) {
                        mbar_wait(&sh->b_empty[stage], sphase ^ 1);
                        if (leader) mbar_expect_tx(&sh->b_full[stage], 2u * (TC2_N / 2) * 128);
                        tma_load_2d_pair(sB + (size_t)stage * (TC2_N / 2) * 128, &tmap_b, mapa_u32(smem_u32(&sh->b_full[stage]), 0),
                                         kb * TC_KB, p * cap + nt * TC2_N + (int)cr * (TC2_N / 2));
                        if (++stage == TC_STAGES) { stage = 0; sphase ^= 1; }
                    }
            }
        }
    } else if (warp == TC_WARP_MMA) {
        // ===================== MMA issuer: one thread of the leader CTA drives both tensor cores =====================
        if (lane == 0 && leader) {
            // D = f32, A and B in `fmt`, K-major, N = 256, M = 256 (128 rows per CTA)
            const uint32_t idesc = (1u << 4) | fmt | ((uint32_t)(TC2_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            const uint64_t a_desc0 = make_sw128_desc(smem_u32(sA)), b_desc0 = make_sw128_desc(smem_u32(sB));
            uint32_t stage = 0, sphase = 0, a_phase = 0, acc = 0, acc_phase = 0;
            for (int item = cluster_id; item < items; item += n_clusters) {
                int p, mt, nt0, nt1;
                if (!item_geom(item, p, mt, nt0, nt1) || nt0 >= nt1) continue;
                mbar_wait(&sh->a_full, a_phase); a_phase ^= 1;
                tc_fence_after();
                for (int nt = nt0; nt < nt1; nt++) {
                    mbar_wait(&sh->acc_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * TC2_N;
                    for (int kb = 0; kb < kblocks; kb++) {
                        mbar_wait(&sh->b_full[stage], sphase);
                        tc_fence_after();
                        const uint64_t a_desc = a_desc0 + (uint64_t)(kb * (TC_M * 128 >> 4));
                        const uint64_t b_desc = b_desc0 + (uint64_t)(stage * ((TC2_N / 2) * 128 >> 4));
                        const int steps = kb == kblocks - 1 ? last_steps : TC_KB / 16;
#pragma unroll
                        for (int k = 0; k < TC_KB / 16; k++)
                            if (k < steps) umma_bf16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) ? 1u : 0u);
                        umma_commit_pair(&sh->b_empty[stage]);
                        if (++stage == TC_STAGES) { stage = 0; sphase ^= 1; }
                    }
                    umma_commit_pair(&sh->acc_full[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                umma_commit_pair(&sh->a_empty);
            }
        }
    } else if (warp < TC_EPI_WARPS) {
        // ===================== epilogue: 32 rows x 128 columns of every 256-wide tile per warp =====================
        const int q4 = warp & 3;
        const int grp = warp >> 2;
        const int row = q4 * 32 + lane;
        uint32_t acc = 0, acc_phase = 0;
        for (int item = cluster_id; item < items; item += n_clusters) {
            int p, mt, nt0, nt1;
            if (!item_geom(item, p, mt, nt0, nt1)) continue;
            float r[4] = { FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX }; int t[4] = { 0, 0, 0, 0 };
            for (int nt = nt0; nt < nt1; nt++) {
                mbar_wait(&sh->acc_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * TC2_N + grp * (TC2_N / TC_EPI_GROUPS);
                const uint32_t acc_empty_leader = mapa_u32(smem_u32(&sh->acc_empty[acc]), 0);
                float l[4] = { FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX };
                float thr = r[3];
                uint32_t v0[32], v1[32];
                static_assert(TC2_N / TC_EPI_GROUPS == 128 || TC2_N / TC_EPI_GROUPS == 64, "two or four column groups");
                if constexpr (TC2_N / TC_EPI_GROUPS == 128) {
                    // TMEM drains at 16 B/clk per sub-partition (256 cycles per x32 load): the next 32 columns are in flight
                    // while the previous 32 go through the selection
                    tmem_ld32_issue(taddr, v0); tmem_ld32_issue(taddr + 32, v1);
                    tmem_ld_wait();
                    tile_top4(v0, 0, l, thr);
                    tmem_ld32_issue(taddr + 64, v0);
                    tile_top4(v1, 32, l, thr);
                    tmem_ld_wait();
                    tmem_ld32_issue(taddr + 96, v1);
                    tile_top4(v0, 64, l, thr);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(acc_empty_leader);            // scores are in registers: hand the buffer back
                    tile_top4(v1, 96, l, thr);
                } else {
                    tmem_ld32_issue(taddr, v0); tmem_ld32_issue(taddr + 32, v1);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(acc_empty_leader);        // scores are in registers: hand the buffer back
                    tile_top4(v0, 0, l, thr); tile_top4(v1, 32, l, thr);
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (!__any_sync(0xffffffffu, l[k] < r[3])) break;
                    running_insert(l[k], nt, r, t);
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            const int q = mt * TC_M + row;
            const int s_id = item % n_splits;
            const size_t o = ((((size_t)p * cap + q) * n_splits + s_id) * TC_EPI_GROUPS + grp) * TC_TOPK;
            epilogue_store(r, t, TC2_N, grp, o, cand_score, cand_idx);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                       // nobody leaves while the pair's MMAs, commits or remote arrives may still land here
    if (warp == TC_WARP_ALLOC) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

// ---------------------------------------------------------------- 3. exact rescoring + guard
// one warp per query.  desc_*: fp32 row-major [img][cap][dim] (the original descriptors).
// G lanes per query (G = candidates per query rounded up to a power of two), 32 / G queries per warp.
// (squared distance, train index) as one ordered 64-bit key: distances are >= 0, so their bit patterns order like the values,
// and equal distances order by index -- the matcher's tie rule.
__device__ __forceinline__ unsigned long long fb_pack(float d, int t) { return ((unsigned long long)__float_as_uint(d) << 32) | (uint32_t)t; }
template <int G>
__global__ void __launch_bounds__(256) rescore_kernel(const float *__restrict__ desc_a, const float *__restrict__ desc_b,
                                                      int64_t pair_stride_a, int64_t pair_stride_b,
                                                      const int32_t *__restrict__ n_a_ptr, int n_a_stride,
                                                      const int32_t *__restrict__ n_b_ptr, int n_b_stride,
                                                      const float *__restrict__ norm_a, const float *__restrict__ norm_b_max,
                                                      const float *__restrict__ err_a, const float *__restrict__ err_b_max,
                                                      const float *__restrict__ cand_score, const int32_t *__restrict__ cand_idx,
                                                      int cap, int dim, int n_splits, int terms, const int32_t *__restrict__ ovf,
                                                      int32_t *best_idx, float *best_dist,
                                                      int32_t *fallback_list, int32_t *fallback_count, int32_t *pair_count,
                                                      unsigned long long *fb_slots)
{
    constexpr int QPW = 32 / G;
    const int p = blockIdx.y;
    const int lane = threadIdx.x & 31, gl = lane % G;
    const int q = (blockIdx.x * 8 + (threadIdx.x >> 5)) * QPW + lane / G;
    const int nA = n_a_ptr[(size_t)p * n_a_stride], nB = n_b_ptr[(size_t)p * n_b_stride];
    const bool live = q < nA;
    const float *a = desc_a + p * pair_stride_a + (size_t)q * dim;
    const float *B = desc_b + p * pair_stride_b;
    const int nc = n_splits * TC_EPI_GROUPS * TC_TOPK;     // <= G candidates: one per lane of the group, lists of TC_TOPK
    float d = FLT_MAX; int t = -1; float approx = FLT_MAX;
    if (live && gl < nc) {
        const size_t o = ((size_t)p * cap + q) * nc + gl;
        t = cand_idx[o]; approx = cand_score[o];
        if (t >= nB || approx > 1e29f) t = -1;      // a padding column (or an unfilled slot): not a candidate, but its key still bounds the list
    }
    // |GEMM score - exact score| <= E_gemm (DESIGN.md "matcher error bound"):
    //   terms 3 (split-bf16): dropped lo*lo and second-order split errors 3*2^-18 |a||b| (x2 for the -2 scale), fp32
    //     accumulation of K' <= 448 exact products with a worst case truncating adder 448*2^-23 (2|a||b| + |b|^2), norm
    //     split 2^-17 |b|^2.
    //   terms 2 / 1 (fp16): the stored query row is a + eps, the stored train row -2b + eta; prep_split_kernel measured
    //     |eps|^2 of this query (err_a) and max |eta|^2 over the pair's train rows (err_b_max) exactly, subnormal rounding
    //     included, so by Cauchy-Schwarz the products are off by at most |a||eta| + 2|eps||b| + |eps||eta|; fp32
    //     accumulation of `used` exact products <= used*2^-23 (2|a||b| + |b|^2); norm: fp32 sum + fp16 split <= 1.1e-6 |b|^2.
    // On top of that a key differs from its score by the column bits (2^-16 |key|), the fp32 rescoring from the real
    // distance by 128*2^-24 d^2, and the fp32 |a|^2 from the real one by < 1e-6 |a|^2: tail(|key|, d^2).
    const float na = live ? norm_a[(size_t)p * cap + q] : 0.f;
    const float bmax = norm_b_max[p];
    const float ab = sqrtf(na) * sqrtf(bmax);
    float E_gemm;
    if (terms == 3) E_gemm = 1.4e-4f * ab + 7e-5f * bmax;
    else {
        const float ea = sqrtf(live ? err_a[(size_t)p * cap + q] : 0.f), eb = sqrtf(err_b_max[p]);
        E_gemm = 1.001f * (sqrtf(na) * eb + 2.f * ea * sqrtf(bmax) + ea * eb) + 1.01f * (float)(terms * dim + 16) * 1.1920929e-7f * (2.f * ab + bmax)
               + 1.1e-6f * bmax + 3e-8f;
    }
#define TC_TAIL(key, dd) (1.6e-5f * fabsf(key) + 8e-6f * (dd) + 1e-6f * na + 1e-7f)
    // Only candidates whose key is within reach of the second smallest key can be one of the best two: a candidate further
    // than 2 E_gemm + the key / rescoring slack above it has a larger fp32 distance than both of the two smallest keys'
    // rows, and its train row is not fetched (typically 2 - 3 of the 8 candidates are).
    {
        const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u) << (lane / G * G));
        const float kv = t >= 0 ? approx : FLT_MAX;
        float k1 = kv;
#pragma unroll
        for (int o = G / 2; o; o >>= 1) k1 = fminf(k1, __shfl_xor_sync(0xffffffffu, k1, o));
        const unsigned holders = __ballot_sync(0xffffffffu, kv == k1) & gmask;
        float k2 = (lane == __ffs(holders) - 1) ? FLT_MAX : kv;
#pragma unroll
        for (int o = G / 2; o; o >>= 1) k2 = fminf(k2, __shfl_xor_sync(0xffffffffu, k2, o));
        const float D = na + bmax + 2.f * ab;
        if (t >= 0 && k2 < 1e29f && approx > k2 + 2.f * E_gemm + 1.6e-5f * (fabsf(approx) + fabsf(k2) + D) + 1e-6f) t = -1;
    }
    {
        if (t >= 0) {
            const float4 *a4 = (const float4 *)a, *b4 = (const float4 *)(B + (size_t)t * dim);     // dim % 4 == 0, rows 16-byte aligned
            float s = 0.f;
#pragma unroll 4
            for (int k = 0; k < dim / 4; k++) {                                                      // serial k order, -fmad=false: the CPU value
                const float4 x = a4[k], y = __ldg(b4 + k);
                float df = x.x - y.x; s += df * df;
                df = x.y - y.y; s += df * df;
                df = x.z - y.z; s += df * df;
                df = x.w - y.w; s += df * df;
            }
            d = s;
        }
    }
    // cmin: the smallest "worst kept" approximate key over the lists that filled up (a list that did not fill saw every
    // train row of its range: it cannot hide a better one)
    float worst = FLT_MAX;
    if (live && gl < nc && (gl % TC_TOPK) == TC_TOPK - 1 && approx < 1e29f) worst = approx;
#pragma unroll
    for (int o = G / 2; o; o >>= 1) worst = fminf(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    // best two by (d, t)
    float d0 = d; int t0 = t;
#pragma unroll
    for (int o = G / 2; o; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, d0, o); const int ot = __shfl_xor_sync(0xffffffffu, t0, o);
        if (ot >= 0 && (t0 < 0 || od < d0 || (od == d0 && ot < t0))) { d0 = od; t0 = ot; }
    }
    float d1 = (t == t0) ? FLT_MAX : d; int t1 = (t == t0) ? -1 : t;
#pragma unroll
    for (int o = G / 2; o; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, d1, o); const int ot = __shfl_xor_sync(0xffffffffu, t1, o);
        if (ot >= 0 && (t1 < 0 || od < d1 || (od == d1 && ot < t1))) { d1 = od; t1 = ot; }
    }
    // every rescored candidate checks the bound it relies on; tests read the counter (fallback_count[1]) and expect 0
    if (t >= 0 && !(fabsf(approx - (d - na)) <= E_gemm + TC_TAIL(approx, d)) && !ovf[p]) atomicAdd(fallback_count + 1, 1);
    if (live && gl == 0) {
        bool ok = !ovf[p];
        if (nB >= 2) {
            if (t1 < 0) ok = false;
            // no train row outside the lists can reach the second best: its key is >= worst, so its fp32 distance is
            // > d1 whenever d1 - |a|^2 stays below worst by the GEMM bound and the tail at (worst, d1)
            else if (worst < FLT_MAX && !((d1 - na) < worst - (E_gemm + TC_TAIL(worst, d1)))) ok = false;
        } else if (nB == 1 && t0 < 0) ok = false;
        const size_t o = ((size_t)p * cap + q) * 2;
        best_idx[o] = t0; best_idx[o + 1] = t1;
        best_dist[o] = t0 >= 0 ? sqrtf(d0) : FLT_MAX; best_dist[o + 1] = t1 >= 0 ? sqrtf(d1) : FLT_MAX;
        if (!ok) {
            // flagged: joins its pair's rescan list, seeded with the candidates' best two so the scan only touches memory for
            // train rows that beat them
            atomicAdd(fallback_count, 1);
            const size_t slot = (size_t)p * cap + atomicAdd(pair_count + p, 1);
            fallback_list[slot] = q;
            fb_slots[2 * slot] = t0 >= 0 ? fb_pack(d0, t0) : ~0ull;
            fb_slots[2 * slot + 1] = t1 >= 0 ? fb_pack(d1, t1) : ~0ull;
        }
    }
}

// per pair: max ||b||^2 and max squared operand rounding error over the train rows
__global__ void norm_max_kernel(const float *__restrict__ norms_b, const float *__restrict__ errs_b, const int32_t *n_b_ptr, int n_b_stride,
                                int cap, float *out, float *err_out)
{
    __shared__ float s[32], se[32];
    const int p = blockIdx.x;
    const int n = n_b_ptr[(size_t)p * n_b_stride];
    float m = 0.f, e = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { m = fmaxf(m, norms_b[(size_t)p * cap + i]); e = fmaxf(e, errs_b[(size_t)p * cap + i]); }
    for (int o = 16; o; o >>= 1) { m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); e = fmaxf(e, __shfl_xor_sync(0xffffffffu, e, o)); }
    if ((threadIdx.x & 31) == 0) { s[threadIdx.x >> 5] = m; se[threadIdx.x >> 5] = e; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) { m = fmaxf(m, s[w]); e = fmaxf(e, se[w]); }
        out[p] = m; err_out[p] = e;
    }
}

// ---------------------------------------------------------------- 4. exact rescan of the flagged queries
// grid (slices of train rows, pairs), 256 threads = 256 train rows at a time.  The flagged queries of a pair are taken
// FB_Q at a time into shared memory and every train row of the slice is read ONCE for all of them (one CTA per query
// streamed the whole train set per query: 4.3 MB of L2 traffic each, 0.3 - 0.6 ms for a few hundred queries).  Distances
// are the serial-k, FMA-free sums of rescore_kernel, so a candidate row met again gives the seed's exact key and is skipped.
// A row whose key beats the query's current second best enters the two-slot list with two atomicMin: slot 0 keeps the
// minimum, every key that loses there (the new one, or the one it displaced) is offered to slot 1 -- the minimum of all
// keys except the overall minimum, in any arrival order.
#define FB_Q 8
// the rows of one slice against the first QN queries in shared memory
template <int QN>
__device__ __forceinline__ void fb_scan_rows(const float *__restrict__ B, int dim, int t_begin, int t_end, const float4 (*s_a)[32],
                                             const unsigned long long *s_seed0, unsigned long long *s_k1, unsigned long long *slots)
{
    const int d4 = dim / 4;
    for (int t = t_begin + threadIdx.x; t < t_end; t += blockDim.x) {
        const float4 *b4 = (const float4 *)(B + (size_t)t * dim);     // dim % 4 == 0, rows 16-byte aligned
        float s[QN];
#pragma unroll
        for (int q = 0; q < QN; q++) s[q] = 0.f;
#pragma unroll 2
        for (int k4 = 0; k4 < d4; k4++) {
            const float4 bv = __ldg(b4 + k4);
#pragma unroll
            for (int q = 0; q < QN; q++) {                              // serial k order per (query, row): the CPU value
                const float4 av = s_a[q][k4];
                float df = av.x - bv.x; s[q] += df * df;
                df = av.y - bv.y; s[q] += df * df;
                df = av.z - bv.z; s[q] += df * df;
                df = av.w - bv.w; s[q] += df * df;
            }
        }
#pragma unroll
        for (int q = 0; q < QN; q++) {
            const unsigned long long key = fb_pack(s[q], t);
            if (key < s_k1[q] && key != s_seed0[q]) {
                unsigned long long *sl = slots + 2 * q;
                const unsigned long long old = atomicMin(sl, key);
                const unsigned long long now1 = min(atomicMin(sl + 1, max(old, key)), max(old, key));
                s_k1[q] = now1;                                          // a looser value written late only costs a few more atomics
            }
        }
    }
}

__global__ void __launch_bounds__(256) fallback_scan_kernel(const float *__restrict__ desc_a, const float *__restrict__ desc_b,
                                                            int64_t pair_stride_a, int64_t pair_stride_b,
                                                            const int32_t *__restrict__ n_b_ptr, int n_b_stride, int cap, int dim,
                                                            const int32_t *__restrict__ pair_count, const int32_t *__restrict__ pair_list,
                                                            unsigned long long *fb_slots)
{
    __shared__ float4 s_a[FB_Q][32];                       // dim <= 128
    __shared__ unsigned long long s_seed0[FB_Q], s_k1[FB_Q];
    const int p = blockIdx.y;
    const int F = pair_count[p];
    if (F == 0) return;
    const int nB = n_b_ptr[(size_t)p * n_b_stride];
    const int per = ((nB + gridDim.x - 1) / gridDim.x + 255) & ~255;          // whole passes of the CTA's 256 threads
    const int t_begin = blockIdx.x * per, t_end = min(nB, t_begin + per);
    if (t_begin >= t_end) return;
    const float *A = desc_a + p * pair_stride_a, *B = desc_b + p * pair_stride_b;
    const int d4 = dim / 4;
    for (int g = 0; g < F; g += FB_Q) {
        const int Q = min(FB_Q, F - g);
        const size_t slot0 = (size_t)p * cap + g;
        __syncthreads();
        for (int i = threadIdx.x; i < FB_Q * d4; i += blockDim.x) {
            const int qi = i / d4, kk = i - qi * d4;
            s_a[qi][kk] = qi < Q ? ((const float4 *)(A + (size_t)pair_list[slot0 + qi] * dim))[kk] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (threadIdx.x < FB_Q) {                          // unused query slots: key 0, nothing is smaller
            s_seed0[threadIdx.x] = (int)threadIdx.x < Q ? fb_slots[2 * (slot0 + threadIdx.x)] : 0ull;
            s_k1[threadIdx.x] = (int)threadIdx.x < Q ? fb_slots[2 * (slot0 + threadIdx.x) + 1] : 0ull;
        }
        __syncthreads();
        if (Q > 4) fb_scan_rows<8>(B, dim, t_begin, t_end, s_a, s_seed0, s_k1, fb_slots + 2 * slot0);
        else if (Q > 2) fb_scan_rows<4>(B, dim, t_begin, t_end, s_a, s_seed0, s_k1, fb_slots + 2 * slot0);
        else if (Q > 1) fb_scan_rows<2>(B, dim, t_begin, t_end, s_a, s_seed0, s_k1, fb_slots + 2 * slot0);
        else fb_scan_rows<1>(B, dim, t_begin, t_end, s_a, s_seed0, s_k1, fb_slots + 2 * slot0);
    }
}

// one thread per flagged query: the two-slot list -> the matcher's output format
__global__ void fallback_store_kernel(const int32_t *__restrict__ pair_count, const int32_t *__restrict__ pair_list,
                                      const unsigned long long *__restrict__ fb_slots, int cap, int32_t *best_idx, float *best_dist)
{
    const int p = blockIdx.y;
    const int F = pair_count[p];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F; i += gridDim.x * blockDim.x) {
        const size_t slot = (size_t)p * cap + i;
        const unsigned long long k0 = fb_slots[2 * slot], k1 = fb_slots[2 * slot + 1];
        const size_t o = ((size_t)p * cap + pair_list[slot]) * 2;
        best_idx[o] = k0 == ~0ull ? -1 : (int32_t)(uint32_t)k0;
        best_idx[o + 1] = k1 == ~0ull ? -1 : (int32_t)(uint32_t)k1;
        best_dist[o] = k0 == ~0ull ? FLT_MAX : sqrtf(__uint_as_float((uint32_t)(k0 >> 32)));
        best_dist[o + 1] = k1 == ~0ull ? FLT_MAX : sqrtf(__uint_as_float((uint32_t)(k1 >> 32)));
    }
}

// ---------------------------------------------------------------- host driver
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
        fn = (PFN_encodeTiled)p;
    }
    return fn;
}

static int make_tmap(CUtensorMap *map, void *base, int kprime, long long rows, bool f16, int box_rows = TC_M)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) { vfsms_set_error("cuTensorMapEncodeTiled not available"); return VFSMS_E_CUDA; }
    cuuint64_t dims[2] = { (cuuint64_t)kprime, (cuuint64_t)rows };
    cuuint64_t strides[1] = { (cuuint64_t)kprime * 2 };
    cuuint32_t box[2] = { TC_KB, (cuuint32_t)box_rows };
    cuuint32_t estr[2] = { 1, 1 };
    CUresult r = enc(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { vfsms_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return VFSMS_E_CUDA; }
    return 0;
}

// desc_a / desc_b: fp32 row-major [pair][cap][dim] (pair stride in floats); counts on device.  Writes best_idx / best_dist
// in the format of match_l2_knn2_batch.
int match_tc_batch(vfsms_ctx *ctx, const float *desc_a, const int32_t *n_a, int n_a_stride, const float *desc_b, const int32_t *n_b,
                   int n_b_stride, int n_pairs, int cap, int dim, int32_t *best_idx, float *best_dist, cudaStream_t st)
{
    if (cap % TC_M || dim % 32 || dim > 128) { vfsms_set_error("match_tc: cap %% 128 == 0 and dim in {32,64,96,128} required"); return VFSMS_E_ARG; }
    MatchWorkspace &mw = ctx->match;
    // operand scheme (vfsms_set_matcher): 3 = split-bf16, 2 = fp16 with the query split in two, 1 = plain fp16
    const int terms = ctx->match_terms;
    const bool f16 = terms < 3;
    const int used = terms * dim + 16;                                  // columns that take part in the product
    const int kblocks = ceil_div(used, TC_KB), kprime = kblocks * TC_KB;
    const int last_steps = (used - (kblocks - 1) * TC_KB) / 16;         // MMA k-steps of the last k-block
    const uint32_t fmt = f16 ? TC_FMT_F16 : TC_FMT_BF16;
    const int m_tiles = cap / TC_M;
    int n_splits = 1;
    while (n_pairs * m_tiles * n_splits < 2 * ctx->num_sms && n_splits * TC_EPI_GROUPS * TC_TOPK < 32) n_splits *= 2;
    int rc;
    const size_t rows = (size_t)n_pairs * cap;
    if ((rc = mw.bf16_a.reserve(rows * kprime * 2))) return rc;
    if ((rc = mw.bf16_b.reserve(rows * kprime * 2))) return rc;
    // cand buffer: scores | idx | norms_a | norms_b | bmax | fallback_count | fallback_list
    const size_t n_cand = rows * n_splits * TC_EPI_GROUPS * TC_TOPK;
    const int np4 = (n_pairs + 3) & ~3;                                 // keeps the float4 scratch below 16-byte aligned
    const size_t bytes = n_cand * 8 + rows * 16 + (size_t)np4 * 16 + 16 + rows * 16 + rows * 4;
    if ((rc = mw.cand_topk.reserve(bytes))) return rc;
    float *cand_score = mw.cand_topk.as<float>();
    int32_t *cand_idx = (int32_t *)(cand_score + n_cand);
    float *norm_a = (float *)(cand_idx + n_cand), *norm_b = norm_a + rows, *err_a = norm_b + rows, *err_b = err_a + rows;
    float *bmax = err_b + rows, *ebmax = bmax + np4;
    int32_t *ovf = (int32_t *)(ebmax + np4);
    int32_t *pair_count = ovf + np4;                                    // flagged queries per pair
    int32_t *fb_count = pair_count + np4;                               // fb_count[0] rescans, [1] error-bound violations
    unsigned long long *fb_slots = (unsigned long long *)(fb_count + 4);    // 8-byte aligned: rows % 128 == 0, np4 % 4 == 0
    int32_t *fb_list = (int32_t *)(fb_slots + 2 * rows);                // [pair][cap] query indices

    StageTimer tt(ctx, st, VFSMS_STAGE_MATCH_TC);
    CUDA_TRY(cudaMemsetAsync(ovf, 0, (size_t)np4 * 8 + 16, st));
    prep_split_kernel<<<dim3(ceil_div(cap, 8), n_pairs), 256, 0, st>>>(desc_a, n_a, n_a_stride, cap, dim, kprime, terms, 0, 0, mw.bf16_a.as<uint16_t>(), norm_a, err_a, ovf);
    LAUNCH_CHECK(ctx);
    prep_split_kernel<<<dim3(ceil_div(cap, 8), n_pairs), 256, 0, st>>>(desc_b, n_b, n_b_stride, cap, dim, kprime, terms, 1, 0, mw.bf16_b.as<uint16_t>(), norm_b, err_b, ovf);
    LAUNCH_CHECK(ctx);
    norm_max_kernel<<<n_pairs, 256, 0, st>>>(norm_b, err_b, n_b, n_b_stride, cap, bmax, ebmax);
    LAUNCH_CHECK(ctx);

    CUtensorMap ta, tb;
    if ((rc = make_tmap(&ta, mw.bf16_a.p, kprime, (long long)rows, f16))) return rc;
    if ((rc = make_tmap(&tb, mw.bf16_b.p, kprime, (long long)rows, f16))) return rc;
    const size_t smem = (size_t)(TC_MAX_KBLOCKS * TC_M + TC_STAGES * TC_N) * 128 + sizeof(TcShared) + 1024;
    static bool attr = false;
    if (!attr) { CUDA_TRY(cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    const int items = n_pairs * m_tiles * n_splits;
    if (ctx->matcher_mode != 2) {
        // default: CTA pairs (cta_group::2), 256 x 256 tiles
        static bool attr_c = false;
        if (!attr_c) { CUDA_TRY(cudaFuncSetAttribute(match_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_c = true; }
        const int citems = n_pairs * ((m_tiles + 1) / 2) * n_splits;
        int clusters = ctx->num_sms / 2;
        if (citems < clusters) clusters = citems;
        match_tc_pair_kernel<<<clusters * 2, TC_THREADS, smem, st>>>(ta, tb, n_a, n_a_stride, n_b, n_b_stride, n_pairs, cap, kblocks, last_steps, fmt, n_splits,
                                                                  cand_score, cand_idx);
        LAUNCH_CHECK(ctx);
    } else {
        const int grid = items < ctx->num_sms ? items : ctx->num_sms;
        match_tc_kernel<<<grid, TC_THREADS, smem, st>>>(ta, tb, n_a, n_a_stride, n_b, n_b_stride, n_pairs, cap, kblocks, last_steps, fmt, n_splits, cand_score, cand_idx);
        LAUNCH_CHECK(ctx);
#ifdef TC_TIMING
        {
            unsigned long long h[8];
            cudaStreamSynchronize(st);
            cudaMemcpyFromSymbol(h, g_tc_t, sizeof(h));
            fprintf(stderr, "[tc timing, sums over %d CTAs] mma: wait_a %llu wait_acc_empty %llu wait_b_full %llu total %llu | epi(w5): wait_acc_full %llu total %llu\n",
                    grid, h[0], h[1], h[2], h[3], h[4], h[5]);
            memset(h, 0, sizeof(h)); cudaMemcpyToSymbol(g_tc_t, h, sizeof(h));
        }
#endif
    }
    {
        const int nc = n_splits * TC_EPI_GROUPS * TC_TOPK;
#define RESCORE(G) rescore_kernel<G><<<dim3(ceil_div(cap, 8 * (32 / G)), n_pairs), 256, 0, st>>>(desc_a, desc_b, (int64_t)cap * dim, (int64_t)cap * dim, \
            n_a, n_a_stride, n_b, n_b_stride, norm_a, bmax, err_a, ebmax, cand_score, cand_idx, cap, dim, n_splits, terms, ovf, best_idx, best_dist, fb_list, fb_count, pair_count, fb_slots)
        if (nc <= 8) RESCORE(8); else if (nc <= 16) RESCORE(16); else RESCORE(32);
#undef RESCORE
        LAUNCH_CHECK(ctx);
    }
    {
        int slices = 4 * ctx->num_sms / n_pairs;
        slices = slices < 1 ? 1 : (slices > 64 ? 64 : slices);
        fallback_scan_kernel<<<dim3(slices, n_pairs), 256, 0, st>>>(desc_a, desc_b, (int64_t)cap * dim, (int64_t)cap * dim, n_b, n_b_stride, cap, dim,
                                                                    pair_count, fb_list, fb_slots);
        LAUNCH_CHECK(ctx);
        fallback_store_kernel<<<dim3(1, n_pairs), 256, 0, st>>>(pair_count, fb_list, fb_slots, cap, best_idx, best_dist);
    }
    LAUNCH_CHECK(ctx);
    ctx->last_fallback_count_dev = fb_count;
    return 0;
}
