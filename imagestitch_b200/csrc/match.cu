// match.cu -- brute-force descriptor matching (exact fp32 path), ratio test and the offset vote, sm_100a.
//
// Replaces Method.matchDescriptors (ImageUtility.py:278-309; plugin side appendix/myGpuFeatures.cpp:148-195) and
// Method.getOffsetByMode (ImageUtility.py:139-178).
//   * knn2_l2_kernel: all-pairs squared L2 with a running top-2 per query.  Every (query, train) distance is
//     accumulated serially over k in fp32 without FMA contraction (this file is built with -fmad=false), i.e. the
//     same value the scalar CPU loop in oracle/surf_oracle.c:so_match_l2_ratio produces; ties -> lower train index.
//   * hamming_best1_kernel: ORB path, byte descriptors stored as float32 (appendix/myGpuFeatures.cpp:118,174-187).
//   * ratio_insert_kernel / vote_reduce_kernel: ratio test, query-ordered compaction, (dRow, dCol) mode with
//     "first seen wins" tie-breaking through a per-pair open-addressing table in global memory -- replaces the
//     O(M^2) Python `list.count` loop that dominates the reference's ORB path.
#include "common.cuh"
#include <float.h>

#define MT 64            // queries / trains per tile
#define MK 32            // k-chunk staged per step
#define M_THREADS 256

// ---------------------------------------------------------------- layout: [n][D] row-major -> k-major [D][ld]
__global__ void transpose_desc_kernel(const float *__restrict__ src, const int32_t *n_ptr, int n_stride, float *dst, int cap, int dim,
                                      int64_t src_pair_stride, int64_t dst_pair_stride)
{
    __shared__ float tile[32][33];
    const int p = blockIdx.z;
    const int n = n_ptr[p * n_stride];
    const float *S = src + p * src_pair_stride;
    float *D = dst + p * dst_pair_stride;
    const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    if (r0 >= n) return;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, k = k0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < n && k < dim) ? S[(size_t)r * dim + k] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int k = k0 + i, r = r0 + threadIdx.x;
        if (k < dim && r < cap) D[(size_t)k * cap + r] = tile[threadIdx.x][i];
    }
}

// ---------------------------------------------------------------- exact fp32 kNN(2)
// grid (ceil(cap/MT), pairs).  descT_*: k-major [dim][cap] per pair.  out: best_idx[q][2], best_dist[q][2] (sqrt).
__global__ void __launch_bounds__(M_THREADS) knn2_l2_kernel(const float *__restrict__ descT_a, const int32_t *n_a_ptr, int n_a_stride,
                                                            const float *__restrict__ descT_b, const int32_t *n_b_ptr, int n_b_stride,
                                                            int cap, int dim, int64_t pair_stride_a, int64_t pair_stride_b,
                                                            int32_t *best_idx, float *best_dist)
{
    __shared__ __align__(16) float As[MK][MT];
    __shared__ __align__(16) float Bs[MK][MT];
    __shared__ float s_d[MT][16][2];
    __shared__ int s_i[MT][16][2];
    const int p = blockIdx.y;
    const int nA = n_a_ptr[p * n_a_stride], nB = n_b_ptr[p * n_b_stride];
    const int q0 = blockIdx.x * MT;
    if (q0 >= nA) return;
    const float *A = descT_a + p * pair_stride_a;
    const float *B = descT_b + p * pair_stride_b;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // thread owns queries ty*4..+3, trains tx*4..+3
    float bd0[4], bd1[4]; int bi0[4], bi1[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { bd0[i] = FLT_MAX; bd1[i] = FLT_MAX; bi0[i] = -1; bi1[i] = -1; }

    for (int t0 = 0; t0 < nB; t0 += MT) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < dim; k0 += MK) {
            __syncthreads();
            for (int idx = threadIdx.x; idx < MK * MT / 4; idx += M_THREADS) {
                const int k = idx / (MT / 4), c = (idx - k * (MT / 4)) * 4;
                float4 va = make_float4(0, 0, 0, 0), vb = make_float4(0, 0, 0, 0);
                if (k0 + k < dim) {
                    va = *(const float4 *)(A + (size_t)(k0 + k) * cap + q0 + c);   // cap is a multiple of 64: in-bounds
                    vb = *(const float4 *)(B + (size_t)(k0 + k) * cap + t0 + c);
                }
                *(float4 *)&As[k][c] = va;
                *(float4 *)&Bs[k][c] = vb;
            }
            __syncthreads();
            const int kmax = min(MK, dim - k0);
#pragma unroll 8
            for (int k = 0; k < kmax; k++) {
                const float4 a = *(const float4 *)&As[k][ty * 4];
                const float4 bq = *(const float4 *)&Bs[k][tx * 4];
                const float av[4] = { a.x, a.y, a.z, a.w }, bv[4] = { bq.x, bq.y, bq.z, bq.w };
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) { const float df = av[i] - bv[j]; acc[i][j] += df * df; }
            }
        }
        // fold the 4 trains of this tile into the running top-2 (ascending train index keeps "first wins")
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int t = t0 + tx * 4 + j;
            if (t < nB) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float s = acc[i][j];
                    if (s < bd0[i]) { bd1[i] = bd0[i]; bi1[i] = bi0[i]; bd0[i] = s; bi0[i] = t; }
                    else if (s < bd1[i]) { bd1[i] = s; bi1[i] = t; }
                }
            }
        }
    }
    // merge the 16 partial top-2 lists of every query; within a thread indices ascend, across threads compare (d, idx)
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) {
        s_d[ty * 4 + i][tx][0] = bd0[i]; s_d[ty * 4 + i][tx][1] = bd1[i];
        s_i[ty * 4 + i][tx][0] = bi0[i]; s_i[ty * 4 + i][tx][1] = bi1[i];
    }
    __syncthreads();
    if (threadIdx.x < MT) {
        const int q = q0 + threadIdx.x;
        if (q < nA) {
            float d0 = FLT_MAX, d1 = FLT_MAX; int i0 = -1, i1 = -1;
            for (int c = 0; c < 16; c++)
                for (int e = 0; e < 2; e++) {
                    const float s = s_d[threadIdx.x][c][e]; const int t = s_i[threadIdx.x][c][e];
                    if (t < 0) continue;
                    // (s, t) lexicographic: equal distance -> lower train index first
                    if (s < d0 || (s == d0 && t < i0)) { d1 = d0; i1 = i0; d0 = s; i0 = t; }
                    else if (s < d1 || (s == d1 && t < i1)) { d1 = s; i1 = t; }
                }
            const size_t o = ((size_t)p * cap + q) * 2;
            best_idx[o] = i0; best_idx[o + 1] = i1;
            best_dist[o] = sqrtf(d0); best_dist[o + 1] = sqrtf(d1);
        }
    }
}

// ---------------------------------------------------------------- Hamming best-1 (ORB): descriptors are bytes stored as float32
// grid (ceil(cap/128), pairs), 128 threads: one query per thread, trains streamed through shared memory.
__global__ void __launch_bounds__(128) hamming_best1_kernel(const float *__restrict__ desc_a, const int32_t *n_a_ptr, int n_a_stride,
                                                            const float *__restrict__ desc_b, const int32_t *n_b_ptr, int n_b_stride,
                                                            int cap, int dim, int64_t pair_stride_a, int64_t pair_stride_b,
                                                            int32_t *best_idx, float *best_dist)
{
    __shared__ uint32_t s_b[128][9];   // 128 trains x up to 32 bytes (8 words, padded)
    const int p = blockIdx.y;
    const int nA = n_a_ptr[p * n_a_stride], nB = n_b_ptr[p * n_b_stride];
    const int q = blockIdx.x * 128 + threadIdx.x;
    if (blockIdx.x * 128 >= nA) return;
    const float *A = desc_a + p * pair_stride_a, *B = desc_b + p * pair_stride_b;
    const int words = dim / 4;     // dim = 32 bytes -> 8 words
    uint32_t qa[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (q < nA)
        for (int w = 0; w < words && w < 8; w++) {
            uint32_t v = 0;
            for (int e = 0; e < 4; e++) v |= ((uint32_t)(int)A[(size_t)q * dim + w * 4 + e] & 255u) << (8 * e);
            qa[w] = v;
        }
    int best = 0x7fffffff, bi = -1;
    for (int t0 = 0; t0 < nB; t0 += 128) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < 128 * words; idx += 128) {
            const int t = idx / words, w = idx - t * words;
            uint32_t v = 0;
            if (t0 + t < nB)
                for (int e = 0; e < 4; e++) v |= ((uint32_t)(int)B[(size_t)(t0 + t) * dim + w * 4 + e] & 255u) << (8 * e);
            s_b[t][w] = v;
        }
        __syncthreads();
        const int tmax = min(128, nB - t0);
        for (int t = 0; t < tmax; t++) {
            int d = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) if (w < words) d += __popc(qa[w] ^ s_b[t][w]);
            if (d < best) { best = d; bi = t0 + t; }
        }
    }
    if (q < nA) {
        const size_t o = ((size_t)p * cap + q) * 2;
        best_idx[o] = bi; best_idx[o + 1] = -1;
        best_dist[o] = (float)best; best_dist[o + 1] = FLT_MAX;
    }
}

// ---------------------------------------------------------------- ratio test + ordered compaction + vote table insert
__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

#define VOTE_EMPTY 0xffffffffu

// mode 0: L2 ratio test (d0 < ratio * d1, needs two neighbours); mode 1: threshold (d0 < param); mode 2: keep all.
// one CTA (1024 threads) per pair.  kp_*: [n][8] records (x, y first).
__global__ void __launch_bounds__(1024) ratio_insert_kernel(const float *__restrict__ kp_a, const float *__restrict__ kp_b,
                                                            int64_t kp_stride_a, int64_t kp_stride_b, int kp_elems,
                                                            const int32_t *n_a_ptr, int n_a_stride,
                                                            const int32_t *__restrict__ best_idx, const float *__restrict__ best_dist,
                                                            int cap, int mode, double param,
                                                            int32_t *matches, int32_t *n_matches,
                                                            uint32_t *tkeys, int32_t *tcnt, int32_t *tfirst, int table_size,
                                                            int do_vote)
{
    __shared__ int s_warp[32];
    __shared__ int s_base, s_total;
    const int p = blockIdx.x;
    const int nA = n_a_ptr[p * n_a_stride];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *KA = kp_a ? kp_a + p * kp_stride_a : nullptr, *KB = kp_b ? kp_b + p * kp_stride_b : nullptr;
    uint32_t *TK = tkeys + (size_t)p * table_size;
    int32_t *TC = tcnt + (size_t)p * table_size, *TF = tfirst + (size_t)p * table_size;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < nA; base += 1024) {
        const int q = base + threadIdx.x;
        bool keep = false; int t = -1;
        if (q < nA) {
            const size_t o = ((size_t)p * cap + q) * 2;
            t = best_idx[o];
            const int t1 = best_idx[o + 1];
            const float d0 = best_dist[o], d1 = best_dist[o + 1];
            if (mode == 0) keep = (t >= 0 && t1 >= 0) && ((double)d0 < (double)d1 * param);
            else if (mode == 1) keep = (t >= 0) && ((double)d0 < param);
            else keep = (t >= 0);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            const int c = s_warp[lane];
            int incl = c;
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            s_warp[lane] = incl - c;
            if (lane == 31) s_total = incl;
        }
        __syncthreads();
        if (keep) {
            const int pos = s_base + s_warp[warp] + __popc(bal & ((1u << lane) - 1));
            matches[((size_t)p * cap + pos) * 2] = t;
            matches[((size_t)p * cap + pos) * 2 + 1] = q;
            if (do_vote) {
                const float ay = KA[(size_t)q * kp_elems + 1], ax = KA[(size_t)q * kp_elems];
                const float by = KB[(size_t)t * kp_elems + 1], bx = KB[(size_t)t * kp_elems];
                const int dr = (int)(ay - by), dc = (int)(ax - bx);     // truncation toward zero (ImageUtility.py:160-161)
                if (dr != 0 || dc != 0) {
                    const uint32_t key = ((uint32_t)(dr + 32768) << 16) | (uint32_t)((dc + 32768) & 0xffff);
                    uint32_t h = hash32(key) & (table_size - 1);
                    while (true) {
                        const uint32_t prev = atomicCAS(&TK[h], VOTE_EMPTY, key);
                        if (prev == VOTE_EMPTY || prev == key) {
                            atomicAdd(&TC[h], 1);
                            atomicMin(&TF[h], pos);
                            break;
                        }
                        h = (h + 1) & (table_size - 1);
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += s_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_matches[p] = s_base;
}

__global__ void __launch_bounds__(1024) vote_reduce_kernel(const uint32_t *__restrict__ tkeys, const int32_t *__restrict__ tcnt,
                                                           const int32_t *__restrict__ tfirst, int table_size,
                                                           const int32_t *n_matches, const int32_t *n_a_ptr, int n_a_stride,
                                                           const int32_t *n_b_ptr, int n_b_stride,
                                                           const int32_t *flags_a, const int32_t *flags_b, int flags_stride,
                                                           int offset_evaluate, vfsms_pair_result *results)
{
    __shared__ int s_cnt[32], s_first[32];
    __shared__ uint32_t s_key[32];
    const int p = blockIdx.x;
    const uint32_t *TK = tkeys + (size_t)p * table_size;
    const int32_t *TC = tcnt + (size_t)p * table_size, *TF = tfirst + (size_t)p * table_size;
    int bc = 0, bf = 0x7fffffff; uint32_t bk = VOTE_EMPTY;
    for (int i = threadIdx.x; i < table_size; i += blockDim.x) {
        const uint32_t k = TK[i];
        if (k == VOTE_EMPTY) continue;
        const int c = TC[i], f = TF[i];
        if (c > bc || (c == bc && f < bf)) { bc = c; bf = f; bk = k; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o; o >>= 1) {
        const int c = __shfl_xor_sync(0xffffffffu, bc, o), f = __shfl_xor_sync(0xffffffffu, bf, o);
        const uint32_t k = __shfl_xor_sync(0xffffffffu, bk, o);
        if (c > bc || (c == bc && f < bf)) { bc = c; bf = f; bk = k; }
    }
    if (lane == 0) { s_cnt[warp] = bc; s_first[warp] = bf; s_key[warp] = bk; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++)
            if (s_cnt[w] > bc || (s_cnt[w] == bc && s_first[w] < bf)) { bc = s_cnt[w]; bf = s_first[w]; bk = s_key[w]; }
        vfsms_pair_result r;
        const int M = n_matches[p];
        r.n_matches = M;
        r.n_a = n_a_ptr ? n_a_ptr[p * n_a_stride] : 0;
        r.n_b = n_b_ptr ? n_b_ptr[p * n_b_stride] : 0;
        r.flags = 0;
        if (flags_a) r.flags |= (flags_a[p * flags_stride] | flags_b[p * flags_stride]) ? 1 : 0;
        if (M == 0) { r.status = 0; r.d_row = 0; r.d_col = 0; r.votes = 0; }
        else if (bc == 0) {     // every match voted exactly (0,0): the reference appends one (0,0) (ImageUtility.py:162-163)
            r.d_row = 0; r.d_col = 0; r.votes = 1; r.status = (1 >= offset_evaluate);
        } else {
            r.d_row = (int)(bk >> 16) - 32768; r.d_col = (int)(bk & 0xffff) - 32768;
            r.votes = bc; r.status = (bc >= offset_evaluate);
        }
        results[p] = r;
    }
}

// ---------------------------------------------------------------- host drivers
static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

int match_reserve(vfsms_ctx *ctx, int n_pairs, int cap)
{
    MatchWorkspace &mw = ctx->match;
    int rc;
    if ((rc = mw.best_idx.reserve((size_t)n_pairs * cap * 2 * 4))) return rc;
    if ((rc = mw.best_dist.reserve((size_t)n_pairs * cap * 2 * 4))) return rc;
    if ((rc = mw.matches.reserve((size_t)n_pairs * cap * 2 * 4))) return rc;
    if ((rc = mw.n_matches.reserve((size_t)n_pairs * 4))) return rc;
    const int ts = next_pow2(cap * 2 < 1024 ? 1024 : cap * 2);
    if ((rc = mw.table_keys.reserve((size_t)n_pairs * ts * 4))) return rc;
    if ((rc = mw.table_cnt.reserve((size_t)n_pairs * ts * 4))) return rc;
    if ((rc = mw.table_first.reserve((size_t)n_pairs * ts * 4))) return rc;
    mw.table_size = ts;
    return 0;
}

int transpose_desc_batch(vfsms_ctx *ctx, const float *src, const int32_t *n_ptr, int n_stride, float *dst, int n_pairs, int cap,
                         int dim, int64_t src_pair_stride, int64_t dst_pair_stride, cudaStream_t st)
{
    dim3 grid(ceil_div(cap, 32), ceil_div(dim, 32), n_pairs), block(32, 8);
    transpose_desc_kernel<<<grid, block, 0, st>>>(src, n_ptr, n_stride, dst, cap, dim, src_pair_stride, dst_pair_stride);
    LAUNCH_CHECK(ctx);
    return 0;
}

// descT_*: k-major [dim][cap] per pair, cap a multiple of 64
int match_l2_knn2_batch(vfsms_ctx *ctx, const float *descT_a, const int32_t *n_a, int n_a_stride,
                        const float *descT_b, const int32_t *n_b, int n_b_stride, int n_pairs, int cap, int dim,
                        int64_t pair_stride_a, int64_t pair_stride_b, int32_t *best_idx, float *best_dist, cudaStream_t st)
{
    if (cap % MT) { vfsms_set_error("match: cap must be a multiple of %d", MT); return VFSMS_E_ARG; }
    knn2_l2_kernel<<<dim3(cap / MT, n_pairs), M_THREADS, 0, st>>>(descT_a, n_a, n_a_stride, descT_b, n_b, n_b_stride, cap, dim,
                                                                  pair_stride_a, pair_stride_b, best_idx, best_dist);
    LAUNCH_CHECK(ctx);
    return 0;
}

int match_hamming_batch(vfsms_ctx *ctx, const float *desc_a, const int32_t *n_a, int n_a_stride,
                        const float *desc_b, const int32_t *n_b, int n_b_stride, int n_pairs, int cap, int dim,
                        int64_t pair_stride_a, int64_t pair_stride_b, int32_t *best_idx, float *best_dist, cudaStream_t st)
{
    if (dim % 4 || dim > 32) { vfsms_set_error("hamming: dim must be a multiple of 4 and <= 32"); return VFSMS_E_ARG; }
    hamming_best1_kernel<<<dim3(ceil_div(cap, 128), n_pairs), 128, 0, st>>>(desc_a, n_a, n_a_stride, desc_b, n_b, n_b_stride, cap, dim,
                                                                            pair_stride_a, pair_stride_b, best_idx, best_dist);
    LAUNCH_CHECK(ctx);
    return 0;
}

// mode: 0 ratio, 1 threshold, 2 keep-all.  kp_* may be null when do_vote == 0.
int ratio_vote_batch(vfsms_ctx *ctx, const float *kp_a, const float *kp_b, int64_t kp_pair_stride_a, int64_t kp_pair_stride_b,
                     int kp_elems, const int32_t *n_a, int n_a_stride, const int32_t *n_b, int n_b_stride,
                     const int32_t *best_idx, const float *best_dist, int n_pairs, int cap, int mode, double param,
                     int offset_evaluate, const int32_t *flags_a, const int32_t *flags_b, int flags_stride, int do_vote,
                     vfsms_pair_result *results, cudaStream_t st)
{
    MatchWorkspace &mw = ctx->match;
    const int ts = mw.table_size;
    if (do_vote) {
        CUDA_TRY(cudaMemsetAsync(mw.table_keys.p, 0xff, (size_t)n_pairs * ts * 4, st));
        CUDA_TRY(cudaMemsetAsync(mw.table_cnt.p, 0, (size_t)n_pairs * ts * 4, st));
        CUDA_TRY(cudaMemsetAsync(mw.table_first.p, 0x7f, (size_t)n_pairs * ts * 4, st));
    }
    ratio_insert_kernel<<<n_pairs, 1024, 0, st>>>(kp_a, kp_b, kp_pair_stride_a, kp_pair_stride_b, kp_elems, n_a, n_a_stride,
                                                  best_idx, best_dist, cap, mode, param, mw.matches.as<int32_t>(),
                                                  mw.n_matches.as<int32_t>(), mw.table_keys.as<uint32_t>(), mw.table_cnt.as<int32_t>(),
                                                  mw.table_first.as<int32_t>(), ts, do_vote);
    LAUNCH_CHECK(ctx);
    if (do_vote) {
        vote_reduce_kernel<<<n_pairs, 1024, 0, st>>>(mw.table_keys.as<uint32_t>(), mw.table_cnt.as<int32_t>(), mw.table_first.as<int32_t>(), ts,
                                                     mw.n_matches.as<int32_t>(), n_a, n_a_stride, n_b, n_b_stride,
                                                     flags_a, flags_b, flags_stride, offset_evaluate, results);
        LAUNCH_CHECK(ctx);
    }
    return 0;
}

// ---------------------------------------------------------------- vote over a caller-provided match list (getOffsetByMode boundary)
__global__ void matches_insert_kernel(const float *__restrict__ kp_a, const float *__restrict__ kp_b, int kp_elems,
                                      const int32_t *__restrict__ matches, int m, int32_t *n_matches,
                                      uint32_t *TK, int32_t *TC, int32_t *TF, int table_size)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) n_matches[0] = m;
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < m; pos += gridDim.x * blockDim.x) {
        const int t = matches[2 * pos], q = matches[2 * pos + 1];
        const float ay = kp_a[(size_t)q * kp_elems + 1], ax = kp_a[(size_t)q * kp_elems];
        const float by = kp_b[(size_t)t * kp_elems + 1], bx = kp_b[(size_t)t * kp_elems];
        const int dr = (int)(ay - by), dc = (int)(ax - bx);
        if (dr == 0 && dc == 0) continue;
        const uint32_t key = ((uint32_t)(dr + 32768) << 16) | (uint32_t)((dc + 32768) & 0xffff);
        uint32_t h = hash32(key) & (table_size - 1);
        while (true) {
            const uint32_t prev = atomicCAS(&TK[h], VOTE_EMPTY, key);
            if (prev == VOTE_EMPTY || prev == key) { atomicAdd(&TC[h], 1); atomicMin(&TF[h], pos); break; }
            h = (h + 1) & (table_size - 1);
        }
    }
}

int vote_matches_batch(vfsms_ctx *ctx, const float *kp_a, const float *kp_b, int kp_elems, const int32_t *matches, int m,
                       int n_a, int n_b, int offset_evaluate, vfsms_pair_result *result_dev, cudaStream_t st)
{
    (void)n_a; (void)n_b;
    MatchWorkspace &mw = ctx->match;
    const int ts = mw.table_size;
    CUDA_TRY(cudaMemsetAsync(mw.table_keys.p, 0xff, (size_t)ts * 4, st));
    CUDA_TRY(cudaMemsetAsync(mw.table_cnt.p, 0, (size_t)ts * 4, st));
    CUDA_TRY(cudaMemsetAsync(mw.table_first.p, 0x7f, (size_t)ts * 4, st));
    const int grid = min(ceil_div(m, 256), ctx->num_sms * 4);
    matches_insert_kernel<<<grid, 256, 0, st>>>(kp_a, kp_b, kp_elems, matches, m, mw.n_matches.as<int32_t>(),
                                                mw.table_keys.as<uint32_t>(), mw.table_cnt.as<int32_t>(), mw.table_first.as<int32_t>(), ts);
    LAUNCH_CHECK(ctx);
    vote_reduce_kernel<<<1, 1024, 0, st>>>(mw.table_keys.as<uint32_t>(), mw.table_cnt.as<int32_t>(), mw.table_first.as<int32_t>(), ts,
                                           mw.n_matches.as<int32_t>(), nullptr, 0, nullptr, 0, nullptr, nullptr, 0, offset_evaluate, result_dev);
    LAUNCH_CHECK(ctx);
    return 0;
}
