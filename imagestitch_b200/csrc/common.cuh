// common.cuh -- shared declarations of libvfsms.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/vfsms.h"

#define KP_X 0
#define KP_Y 1
#define KP_SIZE 2
#define KP_ANGLE 3
#define KP_RESPONSE 4
#define KP_OCTAVE 5
#define KP_LAPLACIAN 6
#define KP_STRIDE VFSMS_KP_STRIDE

#define VFSMS_MAX_OCTAVES 5
#define VFSMS_MAX_LAYERS_PER_OCTAVE 8   // nOctaveLayers + 2
#define VFSMS_NUM_SMS 148

void vfsms_set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            vfsms_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return VFSMS_E_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define LAUNCH_CHECK(ctx)                                                                       \
    do {                                                                                        \
        (ctx)->launches++;                                                                      \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess) {                                                                \
            vfsms_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e));  \
            return VFSMS_E_CUDA;                                                                \
        }                                                                                       \
    } while (0)

// A grow-only device buffer.
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int reserve(size_t n) {
        if (n <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        size_t want = n + n / 8;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { vfsms_set_error("cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e)); return VFSMS_E_CUDA; }
        bytes = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <typename T> T *as() const { return (T *)p; }
};

struct HostBuf {   // pinned staging
    void *p = nullptr;
    size_t bytes = 0;
    int reserve(size_t n) {
        if (n <= bytes) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMallocHost(&p, n + n / 8);
        if (e != cudaSuccess) { vfsms_set_error("cudaMallocHost(%zu) -> %s", n, cudaGetErrorString(e)); return VFSMS_E_CUDA; }
        bytes = n + n / 8;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
    template <typename T> T *as() const { return (T *)p; }
};

// ---------------------------------------------------------------- SURF plan (host-built, passed by value to kernels)
struct HaarBox { short x1, y1, x2, y2; float w; };   // offsets in integral pixels relative to the sample origin

struct SurfLayer {
    int size;            // filter size in pixels
    int margin;          // (size/2)/step : where sample 0 lands in the layer grid
    int samples_i;       // 1 + (rows - size)/step  (0 when the layer is skipped)
    int samples_j;
    HaarBox dx[3], dy[3], dxy[4];
    int off[32];         // the 32 distinct box corners as offsets into the array the octave's kernel reads (pitch folded in)
};

struct SurfPlan {
    int rows, cols;              // image size
    int n_octaves, n_layers;     // n_layers = nOctaveLayers + 2
    float threshold;
    int tile_begin[VFSMS_MAX_OCTAVES + 1];   // prefix of CTA tiles per octave
    int tiles_x[VFSMS_MAX_OCTAVES];
    int stage_off_min[VFSMS_MAX_OCTAVES], stage_rows[VFSMS_MAX_OCTAVES], stage_cols[VFSMS_MAX_OCTAVES];   // smem footprint of a tile
    SurfLayer layer[VFSMS_MAX_OCTAVES][VFSMS_MAX_LAYERS_PER_OCTAVE];
};

// Workspace of the SURF pipeline for a batch of equally-sized images.
struct SurfWorkspace {
    int batch = 0, rows = 0, cols = 0;
    int cand_cap = 0;        // candidate slots per image
    int kp_cap = 0;          // final keypoint slots per image
    int dim = 0;
    int max_features = 0;    // > 0: GPU-plugin semantics (keep the strongest max_features), 0: unlimited
    DevBuf integral;         // [batch][(rows+1)*(cols+1)] int32
    DevBuf band_tot;         // [batch][bands][cols] int32
    DevBuf cand;             // [batch][cand_cap][8] float
    DevBuf sorted;           // [batch][cand_cap][8] float
    DevBuf kp;               // [batch][kp_cap][8] float
    DevBuf desc;             // [batch][kp_cap][dim] float
    DevBuf descT;            // [batch][dim][kp_cap] float, k-major copy for the matcher
    DevBuf counters;         // [batch][4] int32 : n_cand, n_sorted, n_final, flags
    DevBuf prefix;           // [batch+1] int32 prefix of n_final
    DevBuf hist;             // [batch][2048] response histogram + [batch] thresholds
    DevBuf img_f32;          // [batch][rows][pitch_f] float copy of the images * 2^-64 (texture source of the descriptor stage)
    DevBuf fb_list;          // [batch * kp_cap] int32: keypoints the fixed-point sampler hands to the reference sampler
    int pitch_f = 0;
    DevBuf img_off_buf;      // storage of that table (vfsms_tiles_align_list)
    const int64_t *img_off = nullptr;   // device table of per-image byte offsets from base_a for the next surf_run_batch (null: strided layout)
    int last_batch = 0;      // batch of the last surf_run_batch (the work counters sit after its per-image counters)
};

struct MatchWorkspace {
    DevBuf best_idx, best_dist;   // [pairs][cap][2]
    DevBuf matches;               // [pairs][cap][2] int32
    DevBuf n_matches;             // [pairs]
    DevBuf table_keys, table_cnt, table_first;   // vote hash tables [pairs][table_size]
    DevBuf bf16_a, bf16_b;        // tensor-core candidates path
    DevBuf cand_topk;
    int table_size = 0;
};

struct ProfPending { int stage; cudaEvent_t a, b; };

#define MATCH_TERMS_DEFAULT 1
struct vfsms_ctx {
    bool prof = false;
    std::vector<ProfPending> prof_pending;
    std::vector<cudaEvent_t> prof_free;
    float prof_ms[VFSMS_STAGE_COUNT] = {0};
    int32_t prof_calls[VFSMS_STAGE_COUNT] = {0};
    int device = 0;
    int matcher_mode = 0;          // 0: tcgen05 candidates + exact rescoring, 1: exact SIMT kernel, 2: 0 on single CTAs
    int match_terms = MATCH_TERMS_DEFAULT;   // operand scheme of the tcgen05 matcher: 3 split-bf16, 2 fp16 + query split, 1 fp16
    int describe_mode = 1;         // window sampler of the SURF descriptor: 0 reference (double precision, u8 image), 1 fixed-point chunked texture sampler (default)
    int describe_lpt = 1;          // describe the large windows first (two passes over the work list): 0 off, 1 / 2 / 3 = split at 128 / 64 / 256 px
    int sort_mode = 1;             // KeypointGreater ordering: 0 rank by counting over all candidates, 1 rank inside response bins (default)
    int32_t *last_fallback_count_dev = nullptr;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    int num_sms = VFSMS_NUM_SMS;
    SurfWorkspace surf;
    MatchWorkspace match;
    DevBuf img_a, img_b;      // staged ROIs (host variants)
    // double-buffered input slots of vfsms_align_batch_upload / _run: the copy of the next batch runs on copy_stream
    // while the kernels of the current one run on `stream`
    struct UploadSlot { DevBuf a, b; cudaEvent_t uploaded = nullptr; int n_pairs = 0, rows = 0, cols = 0; } slots[2];
    cudaStream_t copy_stream = nullptr;
    DevBuf results;           // vfsms_pair_result[pairs]
    DevBuf scratch0, scratch1, scratch2, scratch3;
    HostBuf pinned_in, pinned_out;
    HostBuf jpeg_pinned;       // entropy-decoded luma coefficients of one chunk of files (jpeg.cu)
    DevBuf jpeg_coef, jpeg_out, jpeg_planes;
    DevBuf tiles;              // device-resident tile stack [tiles_n][tiles_rows][tiles_cols] u8 (vfsms_tiles_*)
    int tiles_n = 0, tiles_rows = 0, tiles_cols = 0;
    bool tiles_borrowed = false;   // `tiles` aliases another context's stack (vfsms_tiles_attach): never freed or resized here
    DevBuf tiles_bgr;          // colour twin of the stack [tiles_n][tiles_rows][tiles_cols][3] (BGR), allocated on first use
    std::vector<uint8_t> tiles_has_bgr;   // per slot: the colour twin holds this tile
    void *tex_cache = nullptr;     // texture objects over caller images (surf.cu)
    DevBuf tex_dev;
    void *phase_state = nullptr;   // cuFFT plans etc. (phase.cu)
    void *blend_state = nullptr;
    void *orb_state = nullptr;
    int entropy_mode = 1;          // JPEG entropy decoding: 0 host threads, 1 on the device (default; vfsms_set_option VFSMS_OPT_ENTROPY)
    int entropy_passes = 0;        // synchronisation passes of the last device entropy decode
    void *jpeg_huff_state = nullptr;  // workspaces of the device entropy decoder (jpeg.cu)
    void *jpeg_enc_state = nullptr;   // coefficient / bit-stream workspaces of the JPEG encoder (jpeg_enc.cu)
};

// stage timing: events on the launching stream; no-ops unless vfsms_profile_enable(ctx, 1)
cudaEvent_t prof_event(vfsms_ctx *ctx);
struct StageTimer {
    vfsms_ctx *ctx; cudaStream_t st; int stage; cudaEvent_t a = nullptr;
    StageTimer(vfsms_ctx *c, cudaStream_t s, int stg) : ctx(c), st(s), stage(stg) {
        if (ctx->prof) { a = prof_event(ctx); cudaEventRecord(a, st); }
    }
    ~StageTimer() {
        if (a) { cudaEvent_t b = prof_event(ctx); cudaEventRecord(b, st); ctx->prof_pending.push_back({stage, a, b}); }
    }
};

// ---------------------------------------------------------------- internal entry points (defined in the .cu files)
int surf_build_plan(SurfPlan *plan, int rows, int cols, const vfsms_surf_params *p);
int surf_reserve(vfsms_ctx *ctx, int batch, int rows, int cols, const vfsms_surf_params *p);
// images: batch images, image b at base_a + b*img_stride for b < split, else base_b + (b-split)*img_stride
int surf_run_batch(vfsms_ctx *ctx, const uint8_t *base_a, const uint8_t *base_b, int split, int batch, int rows,
                   int cols, int stride, int64_t img_stride, const vfsms_surf_params *p, cudaStream_t st);

int surf_grow(vfsms_ctx *ctx, int grow_cand, int grow_kp);
int surf_init_tables();
void surf_tex_destroy(vfsms_ctx *ctx);

// colour decode of n files into out_dev (BGR); gray_out_dev != nullptr also receives the luma plane = what a grayscale decode of the
// same file yields (libjpeg JCS_GRAYSCALE), so one entropy-decoding pass serves alignment (gray) and mosaic (colour)
int jpeg_decode_bgr_gray_dev(vfsms_ctx *ctx, int n, const uint8_t *const *data, const size_t *sizes, uint8_t *out_dev, int rows, int cols,
                             int64_t row_stride, int64_t image_stride, uint8_t *gray_out_dev, int64_t gray_row_stride, int64_t gray_image_stride,
                             cudaStream_t st);
int match_reserve(vfsms_ctx *ctx, int n_pairs, int cap);
int match_tc_batch(vfsms_ctx *ctx, const float *desc_a, const int32_t *n_a, int n_a_stride, const float *desc_b, const int32_t *n_b,
                   int n_b_stride, int n_pairs, int cap, int dim, int32_t *best_idx, float *best_dist, cudaStream_t st);
int transpose_desc_batch(vfsms_ctx *ctx, const float *src, const int32_t *n_ptr, int n_stride, float *dst, int n_pairs, int cap,
                         int dim, int64_t src_pair_stride, int64_t dst_pair_stride, cudaStream_t st);
int match_l2_knn2_batch(vfsms_ctx *ctx, const float *descT_a, const int32_t *n_a, int n_a_stride,
                        const float *descT_b, const int32_t *n_b, int n_b_stride, int n_pairs, int cap, int dim,
                        int64_t pair_stride_a, int64_t pair_stride_b, int32_t *best_idx, float *best_dist, cudaStream_t st);
int match_hamming_batch(vfsms_ctx *ctx, const float *desc_a, const int32_t *n_a, int n_a_stride,
                        const float *desc_b, const int32_t *n_b, int n_b_stride, int n_pairs, int cap, int dim,
                        int64_t pair_stride_a, int64_t pair_stride_b, int32_t *best_idx, float *best_dist, cudaStream_t st);
int ratio_vote_batch(vfsms_ctx *ctx, const float *kp_a, const float *kp_b, int64_t kp_pair_stride_a, int64_t kp_pair_stride_b,
                     int kp_elems, const int32_t *n_a, int n_a_stride, const int32_t *n_b, int n_b_stride,
                     const int32_t *best_idx, const float *best_dist, int n_pairs, int cap, int mode, double param,
                     int offset_evaluate, const int32_t *flags_a, const int32_t *flags_b, int flags_stride, int do_vote,
                     vfsms_pair_result *results, cudaStream_t st);
int vote_matches_batch(vfsms_ctx *ctx, const float *kp_a, const float *kp_b, int kp_elems, const int32_t *matches, int m,
                       int n_a, int n_b, int offset_evaluate, vfsms_pair_result *result_dev, cudaStream_t st);

__host__ __device__ static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
