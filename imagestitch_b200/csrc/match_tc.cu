// match_tc.cu -- tensor-core brute-force matcher (tcgen05 / TMEM / TMA), exact results.  sm_100a only.
//
// Replaces the O(nA*nB*D) part of Method.matchDescriptors (ImageUtility.py:278-309, BF L2 kNN(2)).
//
//   1. prep_split_kernel      fp32 descriptors -> bf16 rows of K' = 3D + 64 columns:
//                               query  row: [ a_hi | a_hi | a_lo | 1, 1, 0 ... ]
//                               train  row: [-2b_hi|-2b_lo|-2b_hi| nb_hi, nb_lo, 0 ... ]     (nb = ||b||^2)
//                             so that  <query row, train row> = ||b||^2 - 2 a.b  up to ~2^-17 relative error
//                             (3-term split-bf16 product; the hi*hi + hi*lo + lo*hi terms of (a_hi+a_lo).(b_hi+b_lo)).
//   2. match_tc_pair_kernel   persistent warp-specialised GEMM on CTA pairs (the default): TMA (SWIZZLE_128B) -> smem ->
//                             tcgen05.mma.cta_group::2 (256 x 256 x 16 per step, kind::f16, bf16 in / fp32 accumulate in TMEM,
//                             two accumulator buffers) -> epilogue warps read TMEM with tcgen05.ld and keep a running top-4
//                             of (score | column) keys per query row and column half, branch-free.  Each CTA's 128 x K'
//                             query tile stays resident in shared memory while its half of every train tile streams through
//                             a 7-stage mbarrier ring.  Nothing but 8 candidates per query is written to HBM.
//      match_tc_kernel        the same on single CTAs (128 x 128 tiles, cta_group::1); kept for verification and as the
//                             subject of the -DTC_TIMING probes that located the shared-memory operand bottleneck.
//   3. rescore_kernel         exact fp32 distances (serial k order, no FMA: the CPU value) of the candidates, best two by
//                             (distance, index), plus a guard: if the second best exact score is not separated from the
//                             worst kept candidate by more than the split-bf16 error bound, the query is flagged ...
//   4. fallback_exact_kernel  ... and rescanned exactly against every train row.  Results are therefore identical to the
//                             exact SIMT kernel (match.cu) by construction, not by luck.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <float.h>

#define TC_M 128
#define TC_N 128
#define TC_KB 64                 // bf16 elements per k-block = 128 bytes = one swizzle row
#define TC_STAGES 7
#define TC_TOPK 4
#define TC_EPI_GROUPS 2          // epilogue warpgroups; group g scans columns [g*64, g*64+64) of every tile
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

// instruction descriptor: D = f32, A = B = bf16, both K-major, N = 128, M = 128
__device__ __forceinline__ uint32_t make_idesc()
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

// ---------------------------------------------------------------- 1. split-bf16 operand rows
// grid (ceil(cap/8), images), 256 threads: one warp per row.  side 0 = query layout, 1 = train layout.
__global__ void __launch_bounds__(256) prep_split_kernel(const float *__restrict__ desc, const int32_t *n_ptr, int n_stride,
                                                         int cap, int dim, int kprime, int side, int img0,
                                                         __nv_bfloat16 *out, float *norms)
{
    const int b = blockIdx.y;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= cap) return;
    const int n = n_ptr[(size_t)(img0 + b) * n_stride];
    const float *src = desc + ((size_t)(img0 + b) * cap + row) * dim;
    __nv_bfloat16 *dst = out + ((size_t)b * cap + row) * kprime;
    const bool valid = row < n;
    float nrm = 0.f;
    for (int k = lane; k < dim; k += 32) {
        const float v = valid ? src[k] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        nrm += v * v;
        if (side == 0) { dst[k] = hi; dst[dim + k] = hi; dst[2 * dim + k] = lo; }
        else {
            const __nv_bfloat16 mhi = __float2bfloat16_rn(-2.f * __bfloat162float(hi)), mlo = __float2bfloat16_rn(-2.f * __bfloat162float(lo));
            dst[k] = mhi; dst[dim + k] = mlo; dst[2 * dim + k] = mhi;
        }
    }
    for (int o = 16; o; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    for (int k = 3 * dim + lane; k < kprime; k += 32) {
        float v = 0.f;
        const int e = k - 3 * dim;
        if (side == 0) v = (e < 2) ? 1.f : 0.f;
        else {
            const float nb = valid ? nrm : 1e30f;       // padding train rows can never be selected
            const float hi = __bfloat162float(__float2bfloat16_rn(nb));
            v = e == 0 ? hi : (e == 1 ? (valid ? nb - hi : 0.f) : 0.f);
        }
        dst[k] = __float2bfloat16_rn(v);
    }
    if (lane == 0) norms[(size_t)b * cap + row] = valid ? nrm : 0.f;
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
__device__ __forceinline__ void tile_top4(const uint32_t (&v)[32], int col_base, float (&l)[4])
{
#pragma unroll
    for (int j = 0; j < 32; j += 4)
        top4_merge4(l, __uint_as_float((v[j] & TC_KEY_MASK) | (uint32_t)(col_base + j)), __uint_as_float((v[j + 1] & TC_KEY_MASK) | (uint32_t)(col_base + j + 1)),
                    __uint_as_float((v[j + 2] & TC_KEY_MASK) | (uint32_t)(col_base + j + 2)), __uint_as_float((v[j + 3] & TC_KEY_MASK) | (uint32_t)(col_base + j + 3)));
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
    static_assert(TC_TOPK == 4 && TC_N / TC_EPI_GROUPS == 64, "epilogue is written for top-4 over 64-column half tiles");
    uint32_t v0[32], v1[32];
    tmem_ld32_issue(taddr, v0);
    tmem_ld32_issue(taddr + 32, v1);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(acc_empty);
    float l[4] = { FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX };
    tile_top4(v0, 0, l);
    tile_top4(v1, 32, l);
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
                                                                 int n_pairs, int cap, int kblocks, int n_splits,
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
            const uint32_t idesc = make_idesc();
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
#pragma unroll
                        for (int k = 0; k < TC_KB / 16; k++)      // +32 B per 16-element k step = +2 in the (addr >> 4) field
                            umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) ? 1u : 0u);
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
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
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
                     int n_pairs, int cap, int kblocks, int n_splits, float *cand_score, int32_t *cand_idx)
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
            // D = f32, A = B = bf16, K-major, N = 256, M = 256 (128 rows per CTA)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC2_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
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
#pragma unroll
                        for (int k = 0; k < TC_KB / 16; k++)
                            umma_bf16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) ? 1u : 0u);
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
                const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * TC2_N + grp * (TC2_N / 2);
                const uint32_t acc_empty_leader = mapa_u32(smem_u32(&sh->acc_empty[acc]), 0);
                float l[4] = { FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX };
                uint32_t v0[32], v1[32];
                tmem_ld32_issue(taddr, v0); tmem_ld32_issue(taddr + 32, v1);
                tmem_ld_wait();
                tile_top4(v0, 0, l); tile_top4(v1, 32, l);
                tmem_ld32_issue(taddr + 64, v0); tmem_ld32_issue(taddr + 96, v1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc_empty_leader);            // scores are in registers: hand the buffer back
                tile_top4(v0, 64, l); tile_top4(v1, 96, l);
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
template <int G>
__global__ void __launch_bounds__(256) rescore_kernel(const float *__restrict__ desc_a, const float *__restrict__ desc_b,
                                                      int64_t pair_stride_a, int64_t pair_stride_b,
                                                      const int32_t *__restrict__ n_a_ptr, int n_a_stride,
                                                      const int32_t *__restrict__ n_b_ptr, int n_b_stride,
                                                      const float *__restrict__ norm_a, const float *__restrict__ norm_b_max,
                                                      const float *__restrict__ cand_score, const int32_t *__restrict__ cand_idx,
                                                      int cap, int dim, int n_splits, int32_t *best_idx, float *best_dist,
                                                      int32_t *fallback_list, int32_t *fallback_count)
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
    if (live && gl == 0) {
        const float na = norm_a[(size_t)p * cap + q];
        const float bmax = norm_b_max[p];
        // |approx key - exact score| <= E.  Terms (DESIGN.md "matcher error bound"): dropped lo*lo and second-order split
        // errors 3*2^-18 |a||b| (x2 for the -2 scale), fp32 accumulation of K' = 448 exact bf16 products, worst case
        // truncating adder 448*2^-23 (2|a||b| + |b|^2), norm split 2^-17 |b|^2, fp32 rescoring 128*2^-24 d^2, and the
        // column bits in the epilogue's keys, 2^-16 |score| <= 2^-16 (|b|^2 + 2|a||b|).
        const float ab = sqrtf(na) * sqrtf(bmax);
        const float E = 1.4e-4f * ab + 7e-5f * bmax + 1e-5f * (na + bmax + 2.f * ab) + 1.6e-5f * (bmax + 2.f * ab) + 1e-7f;
        bool ok = true;
        if (nB >= 2) {
            if (t1 < 0) ok = false;
            else if (worst < FLT_MAX && !((d1 - na) < worst - E)) ok = false;
        } else if (nB == 1 && t0 < 0) ok = false;
        const size_t o = ((size_t)p * cap + q) * 2;
        best_idx[o] = t0; best_idx[o + 1] = t1;
        best_dist[o] = t0 >= 0 ? sqrtf(d0) : FLT_MAX; best_dist[o + 1] = t1 >= 0 ? sqrtf(d1) : FLT_MAX;
        if (!ok) { const int slot = atomicAdd(fallback_count, 1); fallback_list[slot] = p * cap + q; }
    }
}

// per pair: max ||b||^2
__global__ void norm_max_kernel(const float *__restrict__ norms_b, const int32_t *n_b_ptr, int n_b_stride, int cap, float *out)
{
    __shared__ float s[32];
    const int p = blockIdx.x;
    const int n = n_b_ptr[(size_t)p * n_b_stride];
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, norms_b[(size_t)p * cap + i]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) { for (int w = 1; w < (int)(blockDim.x >> 5); w++) m = fmaxf(m, s[w]); out[p] = m; }
}

// ---------------------------------------------------------------- 4. exact fallback for flagged queries (one CTA each)
__device__ __forceinline__ void top2_merge(float &d0, int &t0, float &d1, int &t1, float e0, int u0, float e1, int u1)
{
    float c[4] = { d0, d1, e0, e1 }; int ci[4] = { t0, t1, u0, u1 };
    float b0 = FLT_MAX, b1 = FLT_MAX; int i0 = -1, i1 = -1;
#pragma unroll
    for (int z = 0; z < 4; z++) {
        if (ci[z] < 0) continue;
        if (i0 < 0 || c[z] < b0 || (c[z] == b0 && ci[z] < i0)) { b1 = b0; i1 = i0; b0 = c[z]; i0 = ci[z]; }
        else if (i1 < 0 || c[z] < b1 || (c[z] == b1 && ci[z] < i1)) { b1 = c[z]; i1 = ci[z]; }
    }
    d0 = b0; t0 = i0; d1 = b1; t1 = i1;
}

#define FB_THREADS 512
#define FB_ROWS 4
__global__ void __launch_bounds__(FB_THREADS) fallback_exact_kernel(const float *__restrict__ desc_a, const float *__restrict__ desc_b,
                                                             int64_t pair_stride_a, int64_t pair_stride_b,
                                                             const int32_t *__restrict__ n_b_ptr, int n_b_stride, int cap, int dim,
                                                             const int32_t *__restrict__ fallback_list, const int32_t *__restrict__ fallback_count,
                                                             int32_t *best_idx, float *best_dist)
{
    __shared__ float s_a[128];
    __shared__ float s_d[FB_THREADS / 32][2];
    __shared__ int s_t[FB_THREADS / 32][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = *fallback_count;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int pq = fallback_list[w];
        const int p = pq / cap, q = pq - p * cap;
        const int nB = n_b_ptr[(size_t)p * n_b_stride];
        const float *a = desc_a + p * pair_stride_a + (size_t)q * dim;
        const float *B = desc_b + p * pair_stride_b;
        __syncthreads();
        for (int k = threadIdx.x; k < dim; k += blockDim.x) s_a[k] = a[k];
        __syncthreads();
        float d0 = FLT_MAX, d1 = FLT_MAX; int t0 = -1, t1 = -1;
        // FB_ROWS train rows in flight per thread: the scan is L2-latency bound (few CTAs run), not bandwidth bound
        for (int tb = threadIdx.x; tb < nB; tb += FB_ROWS * FB_THREADS) {
            const float4 *b4[FB_ROWS]; float s[FB_ROWS];
#pragma unroll
            for (int r = 0; r < FB_ROWS; r++) {
                const int t = min(tb + r * FB_THREADS, nB - 1);              // clamped rows are computed and discarded
                b4[r] = (const float4 *)(B + (size_t)t * dim); s[r] = 0.f;   // dim % 4 == 0, rows 16-byte aligned
            }
#pragma unroll 4
            for (int k4 = 0; k4 < dim / 4; k4++) {
                float4 bv[FB_ROWS];
#pragma unroll
                for (int r = 0; r < FB_ROWS; r++) bv[r] = __ldg(b4[r] + k4);
                const float a0 = s_a[4 * k4], a1 = s_a[4 * k4 + 1], a2 = s_a[4 * k4 + 2], a3 = s_a[4 * k4 + 3];
#pragma unroll
                for (int r = 0; r < FB_ROWS; r++) {                          // serial k order per row: the CPU value
                    float df = a0 - bv[r].x; s[r] += df * df;
                    df = a1 - bv[r].y; s[r] += df * df;
                    df = a2 - bv[r].z; s[r] += df * df;
                    df = a3 - bv[r].w; s[r] += df * df;
                }
            }
#pragma unroll
            for (int r = 0; r < FB_ROWS; r++) {                              // ascending t: ties keep the lower train index
                const int t = tb + r * FB_THREADS;
                if (t < nB) {
                    if (s[r] < d0) { d1 = d0; t1 = t0; d0 = s[r]; t0 = t; }
                    else if (s[r] < d1) { d1 = s[r]; t1 = t; }
                }
            }
        }
        for (int o = 16; o; o >>= 1) {
            const float e0 = __shfl_xor_sync(0xffffffffu, d0, o), e1 = __shfl_xor_sync(0xffffffffu, d1, o);
            const int u0 = __shfl_xor_sync(0xffffffffu, t0, o), u1 = __shfl_xor_sync(0xffffffffu, t1, o);
            top2_merge(d0, t0, d1, t1, e0, u0, e1, u1);
        }
        if (lane == 0) { s_d[warp][0] = d0; s_d[warp][1] = d1; s_t[warp][0] = t0; s_t[warp][1] = t1; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int z = 1; z < FB_THREADS / 32; z++) top2_merge(d0, t0, d1, t1, s_d[z][0], s_t[z][0], s_d[z][1], s_t[z][1]);
            const size_t o = ((size_t)p * cap + q) * 2;
            best_idx[o] = t0; best_idx[o + 1] = t1;
            best_dist[o] = t0 >= 0 ? sqrtf(d0) : FLT_MAX; best_dist[o + 1] = t1 >= 0 ? sqrtf(d1) : FLT_MAX;
        }
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

static int make_tmap(CUtensorMap *map, void *base, int kprime, long long rows, int box_rows = TC_M)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) { vfsms_set_error("cuTensorMapEncodeTiled not available"); return VFSMS_E_CUDA; }
    cuuint64_t dims[2] = { (cuuint64_t)kprime, (cuuint64_t)rows };
    cuuint64_t strides[1] = { (cuuint64_t)kprime * 2 };
    cuuint32_t box[2] = { TC_KB, (cuuint32_t)box_rows };
    cuuint32_t estr[2] = { 1, 1 };
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
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
    const int kprime = 3 * dim + 64, kblocks = kprime / TC_KB;
    const int m_tiles = cap / TC_M;
    int n_splits = 1;
    while (n_pairs * m_tiles * n_splits < 2 * ctx->num_sms && n_splits * TC_EPI_GROUPS * TC_TOPK < 32) n_splits *= 2;
    int rc;
    const size_t rows = (size_t)n_pairs * cap;
    if ((rc = mw.bf16_a.reserve(rows * kprime * 2))) return rc;
    if ((rc = mw.bf16_b.reserve(rows * kprime * 2))) return rc;
    // cand buffer: scores | idx | norms_a | norms_b | bmax | fallback_count | fallback_list
    const size_t n_cand = rows * n_splits * TC_EPI_GROUPS * TC_TOPK;
    const size_t bytes = n_cand * 8 + rows * 8 + (size_t)n_pairs * 4 + 16 + rows * 4;
    if ((rc = mw.cand_topk.reserve(bytes))) return rc;
    float *cand_score = mw.cand_topk.as<float>();
    int32_t *cand_idx = (int32_t *)(cand_score + n_cand);
    float *norm_a = (float *)(cand_idx + n_cand), *norm_b = norm_a + rows, *bmax = norm_b + rows;
    int32_t *fb_count = (int32_t *)(bmax + n_pairs), *fb_list = fb_count + 4;

    StageTimer tt(ctx, st, VFSMS_STAGE_MATCH_TC);
    prep_split_kernel<<<dim3(ceil_div(cap, 8), n_pairs), 256, 0, st>>>(desc_a, n_a, n_a_stride, cap, dim, kprime, 0, 0, mw.bf16_a.as<__nv_bfloat16>(), norm_a);
    LAUNCH_CHECK(ctx);
    prep_split_kernel<<<dim3(ceil_div(cap, 8), n_pairs), 256, 0, st>>>(desc_b, n_b, n_b_stride, cap, dim, kprime, 1, 0, mw.bf16_b.as<__nv_bfloat16>(), norm_b);
    LAUNCH_CHECK(ctx);
    norm_max_kernel<<<n_pairs, 256, 0, st>>>(norm_b, n_b, n_b_stride, cap, bmax);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(cudaMemsetAsync(fb_count, 0, 16, st));

    CUtensorMap ta, tb;
    if ((rc = make_tmap(&ta, mw.bf16_a.p, kprime, (long long)rows))) return rc;
    if ((rc = make_tmap(&tb, mw.bf16_b.p, kprime, (long long)rows))) return rc;
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
        match_tc_pair_kernel<<<clusters * 2, TC_THREADS, smem, st>>>(ta, tb, n_a, n_a_stride, n_b, n_b_stride, n_pairs, cap, kblocks, n_splits,
                                                                  cand_score, cand_idx);
        LAUNCH_CHECK(ctx);
    } else {
        const int grid = items < ctx->num_sms ? items : ctx->num_sms;
        match_tc_kernel<<<grid, TC_THREADS, smem, st>>>(ta, tb, n_a, n_a_stride, n_b, n_b_stride, n_pairs, cap, kblocks, n_splits, cand_score, cand_idx);
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
            n_a, n_a_stride, n_b, n_b_stride, norm_a, bmax, cand_score, cand_idx, cap, dim, n_splits, best_idx, best_dist, fb_list, fb_count)
        if (nc <= 8) RESCORE(8); else if (nc <= 16) RESCORE(16); else RESCORE(32);
#undef RESCORE
        LAUNCH_CHECK(ctx);
    }
    fallback_exact_kernel<<<ctx->num_sms * 2, FB_THREADS, 0, st>>>(desc_a, desc_b, (int64_t)cap * dim, (int64_t)cap * dim, n_b, n_b_stride, cap, dim,
                                                            fb_list, fb_count, best_idx, best_dist);
    LAUNCH_CHECK(ctx);
    ctx->last_fallback_count_dev = fb_count;
    return 0;
}
