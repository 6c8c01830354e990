// capi.cu -- extern "C" surface of libvfsms.so (see include/vfsms.h for the contract and reference citations).
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <algorithm>
#include <vector>

static thread_local char g_err[1024] = "";

void vfsms_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// phase.cu / blend.cu / orb.cu
void phase_state_destroy(vfsms_ctx *ctx);
void blend_state_destroy(vfsms_ctx *ctx);
void orb_state_destroy(vfsms_ctx *ctx);
void jpeg_enc_state_destroy(vfsms_ctx *ctx);
void jpeg_huff_state_destroy(vfsms_ctx *ctx);

cudaEvent_t prof_event(vfsms_ctx *ctx)
{
    if (!ctx->prof_free.empty()) { cudaEvent_t e = ctx->prof_free.back(); ctx->prof_free.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}

static const char *k_stage_names[VFSMS_STAGE_COUNT] = { "integral", "hessian_nms", "rank_sort", "validate_compact", "orient_describe",
                                                        "transpose", "match_knn2", "ratio_vote", "phase_fft", "phase_peak", "blend",
                                                        "match_tc" };

extern "C" {

int vfsms_profile_enable(vfsms_ctx *ctx, int on) { if (!ctx) return VFSMS_E_ARG; ctx->prof = on != 0; return 0; }
const char *vfsms_stage_name(int stage) { return (stage >= 0 && stage < VFSMS_STAGE_COUNT) ? k_stage_names[stage] : ""; }
int vfsms_profile_read(vfsms_ctx *ctx, float *ms_out, int32_t *calls_out, int reset)
{
    if (!ctx) return VFSMS_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaDeviceSynchronize());
    for (auto &pp : ctx->prof_pending) {
        float ms = 0; cudaEventElapsedTime(&ms, pp.a, pp.b);
        ctx->prof_ms[pp.stage] += ms; ctx->prof_calls[pp.stage]++;
        ctx->prof_free.push_back(pp.a); ctx->prof_free.push_back(pp.b);
    }
    ctx->prof_pending.clear();
    for (int i = 0; i < VFSMS_STAGE_COUNT; i++) {
        if (ms_out) ms_out[i] = ctx->prof_ms[i];
        if (calls_out) calls_out[i] = ctx->prof_calls[i];
        if (reset) { ctx->prof_ms[i] = 0; ctx->prof_calls[i] = 0; }
    }
    return 0;
}

int vfsms_set_matcher(vfsms_ctx *ctx, int mode)
{
    if (!ctx || mode < 0 || mode > 5) { vfsms_set_error("vfsms_set_matcher: bad arguments"); return VFSMS_E_ARG; }
    ctx->matcher_mode = mode >= 3 ? 0 : mode;
    ctx->match_terms = mode >= 3 ? 6 - mode : MATCH_TERMS_DEFAULT;      // 3 -> split-bf16, 4 -> fp16 two terms, 5 -> fp16 one term
    return 0;
}
static const char *const k_option_names[VFSMS_OPT_COUNT] = { "describe", "sort", "lpt", "entropy" };
static const int k_option_max[VFSMS_OPT_COUNT] = { 3, 1, 3, 1 };
static int *option_slot(vfsms_ctx *ctx, int option)
{
    switch (option) {
    case VFSMS_OPT_DESCRIBE_MODE: return &ctx->describe_mode;
    case VFSMS_OPT_SORT_MODE: return &ctx->sort_mode;
    case VFSMS_OPT_DESCRIBE_LPT: return &ctx->describe_lpt;
    case VFSMS_OPT_ENTROPY: return &ctx->entropy_mode;
    default: return nullptr;
    }
}
const char *vfsms_option_name(int option) { return option >= 0 && option < VFSMS_OPT_COUNT ? k_option_names[option] : "?"; }
int vfsms_set_option(vfsms_ctx *ctx, int option, int value)
{
    int *slot = ctx ? option_slot(ctx, option) : nullptr;
    if (!slot || value < 0 || value > k_option_max[option]) { vfsms_set_error("vfsms_set_option: bad option %d / value %d", option, value); return VFSMS_E_ARG; }
    *slot = value;
    return 0;
}
int vfsms_get_option(vfsms_ctx *ctx, int option, int *value_out)
{
    int *slot = ctx ? option_slot(ctx, option) : nullptr;
    if (!slot || !value_out) { vfsms_set_error("vfsms_get_option: bad arguments"); return VFSMS_E_ARG; }
    *value_out = *slot;
    return 0;
}
// VFSMS_OPTS="describe=2,sort=1": applied by vfsms_create; a malformed entry fails the create (no silent defaults)
static int apply_env_options(vfsms_ctx *ctx)
{
    const char *env = getenv("VFSMS_OPTS");
    if (!env || !*env) return 0;
    std::string all(env);
    size_t pos = 0;
    while (pos <= all.size()) {
        size_t end = all.find(',', pos);
        if (end == std::string::npos) end = all.size();
        const std::string item = all.substr(pos, end - pos);
        pos = end + 1;
        if (item.empty()) continue;
        const size_t eq = item.find('=');
        int opt = -1;
        if (eq != std::string::npos)
            for (int o = 0; o < VFSMS_OPT_COUNT; o++) if (item.compare(0, eq, k_option_names[o]) == 0) opt = o;
        if (opt < 0) { vfsms_set_error("VFSMS_OPTS: cannot parse '%s'", item.c_str()); return VFSMS_E_ARG; }
        int rc = vfsms_set_option(ctx, opt, atoi(item.c_str() + eq + 1));
        if (rc) return rc;
    }
    return 0;
}
int vfsms_last_match_fallbacks(vfsms_ctx *ctx, int *count_out)
{
    if (!ctx || !count_out) return VFSMS_E_ARG;
    *count_out = 0;
    if (!ctx->last_fallback_count_dev) return 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(count_out, ctx->last_fallback_count_dev, 4, cudaMemcpyDeviceToHost));
    return 0;
}
int vfsms_last_match_bound_violations(vfsms_ctx *ctx, int *count_out)
{
    if (!ctx || !count_out) return VFSMS_E_ARG;
    *count_out = 0;
    if (!ctx->last_fallback_count_dev) return 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(count_out, ctx->last_fallback_count_dev + 1, 4, cudaMemcpyDeviceToHost));
    return 0;
}

int vfsms_last_describe_handovers(vfsms_ctx *ctx, int *count_out)
{
    if (!ctx || !count_out) return VFSMS_E_ARG;
    *count_out = 0;
    if (!ctx->surf.counters.p || ctx->surf.batch <= 0) return 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(count_out, ctx->surf.counters.as<int32_t>() + (size_t)ctx->surf.last_batch * 4 + 2, 4, cudaMemcpyDeviceToHost));
    return 0;
}

int vfsms_version(void) { return VFSMS_VERSION; }
const char *vfsms_last_error(void) { return g_err; }

int vfsms_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int vfsms_create(int device, vfsms_ctx **out)
{
    if (!out) { vfsms_set_error("vfsms_create: out is NULL"); return VFSMS_E_ARG; }
    *out = nullptr;
    int n = vfsms_device_count();
    if (n <= 0) { vfsms_set_error("no CUDA device visible: libvfsms has no CPU fallback"); return VFSMS_E_NODEVICE; }
    if (device < 0 || device >= n) { vfsms_set_error("device %d out of range (0..%d)", device, n - 1); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { vfsms_set_error("device %d is sm_%d%d; libvfsms is built for sm_100a only", device, prop.major, prop.minor); return VFSMS_E_NODEVICE; }
    vfsms_ctx *ctx = new vfsms_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    int rc = surf_init_tables();
    if (rc) { delete ctx; return rc; }
    if ((rc = apply_env_options(ctx))) { cudaStreamDestroy(ctx->stream); delete ctx; return rc; }
    *out = ctx;
    return 0;
}

void vfsms_destroy(vfsms_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &pp : ctx->prof_pending) { cudaEventDestroy(pp.a); cudaEventDestroy(pp.b); }
    for (auto e : ctx->prof_free) cudaEventDestroy(e);
    phase_state_destroy(ctx);
    blend_state_destroy(ctx);
    orb_state_destroy(ctx);
    jpeg_enc_state_destroy(ctx);
    jpeg_huff_state_destroy(ctx);
    surf_tex_destroy(ctx);
    ctx->tex_dev.release();
    SurfWorkspace &w = ctx->surf;
    DevBuf *bufs[] = { &w.integral, &w.band_tot, &w.cand, &w.sorted, &w.kp, &w.desc, &w.descT, &w.counters, &w.prefix, &w.hist, &w.img_f32, &w.fb_list, &w.img_off_buf,
                       &ctx->match.best_idx, &ctx->match.best_dist, &ctx->match.matches, &ctx->match.n_matches,
                       &ctx->match.table_keys, &ctx->match.table_cnt, &ctx->match.table_first, &ctx->match.bf16_a,
                       &ctx->match.bf16_b, &ctx->match.cand_topk, &ctx->img_a, &ctx->img_b, &ctx->results,
                       &ctx->scratch0, &ctx->scratch1, &ctx->scratch2, &ctx->scratch3 };
    for (DevBuf *b : bufs) b->release();
    ctx->pinned_in.release(); ctx->pinned_out.release();
    for (auto &sl : ctx->slots) { sl.a.release(); sl.b.release(); if (sl.uploaded) cudaEventDestroy(sl.uploaded); }
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->tiles_borrowed) { ctx->tiles.p = nullptr; ctx->tiles.bytes = 0; }
    ctx->jpeg_pinned.release(); ctx->jpeg_coef.release(); ctx->jpeg_out.release(); ctx->jpeg_planes.release(); ctx->tiles.release(); ctx->tiles_bgr.release();
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int vfsms_synchronize(vfsms_ctx *ctx)
{
    if (!ctx) return VFSMS_E_ARG;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

void *vfsms_stream(vfsms_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int64_t vfsms_launch_count(vfsms_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------- SURF, host in / host out
int vfsms_surf_detect_and_describe(vfsms_ctx *ctx, const uint8_t *image, int rows, int cols, int stride,
                                   const vfsms_surf_params *params, float *kp_out, float *desc_out, int cap, int *n_out)
{
    if (!ctx || !image || !params || !n_out || rows < 1 || cols < 1 || stride < cols) { vfsms_set_error("surf: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = ctx->img_a.reserve((size_t)rows * cols))) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(ctx->img_a.p, cols, image, stride, cols, rows, cudaMemcpyHostToDevice, st));
    if ((rc = surf_reserve(ctx, 1, rows, cols, params))) return rc;
    int32_t cnt[4];
    for (int attempt = 0; attempt < 6; attempt++) {
        if ((rc = surf_run_batch(ctx, ctx->img_a.as<uint8_t>(), nullptr, 1, 1, rows, cols, cols, 0, params, st))) return rc;
        CUDA_TRY(cudaMemcpyAsync(cnt, ctx->surf.counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (!(cnt[3] & 3)) break;
        if ((rc = surf_grow(ctx, cnt[3] & 1, cnt[3] & 2))) return rc;
        if (attempt == 5) { vfsms_set_error("surf: candidate buffer overflow after regrow"); return VFSMS_E_OVERFLOW; }
    }
    const int n = cnt[2];
    *n_out = n;
    if (n > cap) { vfsms_set_error("surf: %d keypoints but capacity %d", n, cap); return VFSMS_E_CAPACITY; }
    const int dim = ctx->surf.dim;
    if (n > 0) {
        if (kp_out) CUDA_TRY(cudaMemcpyAsync(kp_out, ctx->surf.kp.p, (size_t)n * KP_STRIDE * 4, cudaMemcpyDeviceToHost, st));
        if (desc_out) CUDA_TRY(cudaMemcpyAsync(desc_out, ctx->surf.desc.p, (size_t)n * dim * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return 0;
}

// ---------------------------------------------------------------- matcher, host in / host out
int vfsms_match_descriptors(vfsms_ctx *ctx, const float *desc_a, int n_a, const float *desc_b, int n_b, int dim,
                            int feature_type, float param, int32_t *matches_out, int *m_out)
{
    if (!ctx || !m_out || n_a < 0 || n_b < 0 || dim < 1) { vfsms_set_error("match: bad arguments"); return VFSMS_E_ARG; }
    if (feature_type != 1 && feature_type != 2 && feature_type != 3) { vfsms_set_error("match: featureType %d", feature_type); return VFSMS_E_ARG; }
    *m_out = 0;
    if (n_a == 0 || n_b == 0) return 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int cap = (((n_a > n_b ? n_a : n_b) + 255) / 256) * 256;
    int rc;
    if ((rc = ctx->scratch0.reserve((size_t)cap * dim * 4))) return rc;   // A row-major
    if ((rc = ctx->scratch1.reserve((size_t)cap * dim * 4))) return rc;   // B row-major
    if ((rc = ctx->scratch2.reserve((size_t)cap * dim * 4 * 2))) return rc;   // A^T, B^T
    if ((rc = ctx->scratch3.reserve(64))) return rc;                      // counts
    if ((rc = match_reserve(ctx, 1, cap))) return rc;
    int32_t counts[2] = { n_a, n_b };
    CUDA_TRY(cudaMemcpyAsync(ctx->scratch3.p, counts, sizeof(counts), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ctx->scratch0.p, desc_a, (size_t)n_a * dim * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ctx->scratch1.p, desc_b, (size_t)n_b * dim * 4, cudaMemcpyHostToDevice, st));
    const int32_t *cn = ctx->scratch3.as<int32_t>();
    MatchWorkspace &mw = ctx->match;
    if (feature_type == 3) {
        if ((rc = match_hamming_batch(ctx, ctx->scratch0.as<float>(), cn, 0, ctx->scratch1.as<float>(), cn + 1, 0, 1, cap, dim, 0, 0,
                                      mw.best_idx.as<int32_t>(), mw.best_dist.as<float>(), st))) return rc;
    } else if (ctx->matcher_mode != 1 && dim % 32 == 0 && dim <= 128) {
        // descriptors of unknown scale (SIFT: values up to 255) take the split-bf16 operands: fp16 has no range for them
        const int terms_saved = ctx->match_terms;
        if (feature_type != 2) ctx->match_terms = 3;
        rc = match_tc_batch(ctx, ctx->scratch0.as<float>(), cn, 0, ctx->scratch1.as<float>(), cn + 1, 0, 1, cap, dim,
                            mw.best_idx.as<int32_t>(), mw.best_dist.as<float>(), st);
        ctx->match_terms = terms_saved;
        if (rc) return rc;
    } else {
        float *AT = ctx->scratch2.as<float>(), *BT = AT + (size_t)cap * dim;
        CUDA_TRY(cudaMemsetAsync(AT, 0, (size_t)cap * dim * 4 * 2, st));
        if ((rc = transpose_desc_batch(ctx, ctx->scratch0.as<float>(), cn, 0, AT, 1, cap, dim, 0, 0, st))) return rc;
        if ((rc = transpose_desc_batch(ctx, ctx->scratch1.as<float>(), cn + 1, 0, BT, 1, cap, dim, 0, 0, st))) return rc;
        if ((rc = match_l2_knn2_batch(ctx, AT, cn, 0, BT, cn + 1, 0, 1, cap, dim, 0, 0, mw.best_idx.as<int32_t>(), mw.best_dist.as<float>(), st))) return rc;
    }
    if ((rc = ratio_vote_batch(ctx, nullptr, nullptr, 0, 0, KP_STRIDE, cn, 0, cn + 1, 0, mw.best_idx.as<int32_t>(), mw.best_dist.as<float>(),
                               1, cap, feature_type == 3 ? 1 : 0, (double)param, 0, nullptr, nullptr, 0, 0, nullptr, st))) return rc;
    int32_t m = 0;
    CUDA_TRY(cudaMemcpyAsync(&m, mw.n_matches.p, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *m_out = m;
    if (m > 0 && matches_out) {
        CUDA_TRY(cudaMemcpyAsync(matches_out, mw.matches.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return 0;
}

// ---------------------------------------------------------------- getOffsetByMode
int vfsms_offset_by_mode(vfsms_ctx *ctx, const float *kps_a, int n_a, const float *kps_b, int n_b, int kp_stride,
                         const int32_t *matches, int m, int offset_evaluate, vfsms_pair_result *result)
{
    if (!ctx || !result || m < 0 || kp_stride < 2) { vfsms_set_error("offset_by_mode: bad arguments"); return VFSMS_E_ARG; }
    memset(result, 0, sizeof(*result));
    result->n_a = n_a; result->n_b = n_b; result->n_matches = m;
    if (m == 0) return 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = ctx->scratch0.reserve((size_t)n_a * kp_stride * 4))) return rc;
    if ((rc = ctx->scratch1.reserve((size_t)n_b * kp_stride * 4))) return rc;
    if ((rc = ctx->scratch2.reserve((size_t)m * 8))) return rc;
    if ((rc = ctx->results.reserve(sizeof(vfsms_pair_result)))) return rc;
    if ((rc = match_reserve(ctx, 1, m))) return rc;
    CUDA_TRY(cudaMemcpyAsync(ctx->scratch0.p, kps_a, (size_t)n_a * kp_stride * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ctx->scratch1.p, kps_b, (size_t)n_b * kp_stride * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ctx->scratch2.p, matches, (size_t)m * 8, cudaMemcpyHostToDevice, st));
    if ((rc = vote_matches_batch(ctx, ctx->scratch0.as<float>(), ctx->scratch1.as<float>(), kp_stride, ctx->scratch2.as<int32_t>(), m,
                                 n_a, n_b, offset_evaluate, ctx->results.as<vfsms_pair_result>(), st))) return rc;
    CUDA_TRY(cudaMemcpyAsync(result, ctx->results.p, sizeof(*result), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    result->n_a = n_a; result->n_b = n_b;
    return 0;
}

// ---------------------------------------------------------------- fused batch alignment
static int align_batch_launch(vfsms_ctx *ctx, const uint8_t *a_dev, const uint8_t *b_dev, int n_pairs, int rows, int cols,
                              int stride, int64_t pair_stride, const vfsms_surf_params *params, double ratio,
                              int offset_evaluate, vfsms_pair_result *results_dev, cudaStream_t st)
{
    int rc;
    SurfWorkspace &ws = ctx->surf;
    if ((rc = surf_run_batch(ctx, a_dev, b_dev, n_pairs, 2 * n_pairs, rows, cols, stride, pair_stride, params, st))) return rc;
    const int cap = ws.kp_cap, dim = ws.dim;
    const int32_t *nfin = ws.counters.as<int32_t>() + 2;     // n_final of image b at [b*4]
    const int32_t *flags = ws.counters.as<int32_t>() + 3;
    MatchWorkspace &mw = ctx->match;
    const bool use_tc = ctx->matcher_mode != 1 && cap % 128 == 0 && dim % 32 == 0 && dim <= 128;
    if (use_tc) {
        const float *DA = ws.desc.as<float>(), *DB = DA + (size_t)n_pairs * cap * dim;
        if ((rc = match_tc_batch(ctx, DA, nfin, 4, DB, nfin + 4 * n_pairs, 4, n_pairs, cap, dim, mw.best_idx.as<int32_t>(),
                                 mw.best_dist.as<float>(), st))) return rc;
    } else {
        {
            StageTimer t(ctx, st, VFSMS_STAGE_TRANSPOSE);
            CUDA_TRY(cudaMemsetAsync(ws.descT.p, 0, (size_t)2 * n_pairs * cap * dim * 4, st));
            if ((rc = transpose_desc_batch(ctx, ws.desc.as<float>(), nfin, 4, ws.descT.as<float>(), 2 * n_pairs, cap, dim,
                                           (int64_t)cap * dim, (int64_t)cap * dim, st))) return rc;
        }
        const float *AT = ws.descT.as<float>(), *BT = AT + (size_t)n_pairs * cap * dim;
        StageTimer t(ctx, st, VFSMS_STAGE_MATCH);
        if ((rc = match_l2_knn2_batch(ctx, AT, nfin, 4, BT, nfin + 4 * n_pairs, 4, n_pairs, cap, dim, (int64_t)cap * dim, (int64_t)cap * dim,
                                      mw.best_idx.as<int32_t>(), mw.best_dist.as<float>(), st))) return rc;
    }
    const float *KA = ws.kp.as<float>(), *KB = KA + (size_t)n_pairs * cap * KP_STRIDE;
    StageTimer tv(ctx, st, VFSMS_STAGE_VOTE);
    if ((rc = ratio_vote_batch(ctx, KA, KB, (int64_t)cap * KP_STRIDE, (int64_t)cap * KP_STRIDE, KP_STRIDE, nfin, 4, nfin + 4 * n_pairs, 4,
                               mw.best_idx.as<int32_t>(), mw.best_dist.as<float>(), n_pairs, cap, 0, ratio, offset_evaluate,
                               flags, flags + 4 * n_pairs, 4, 1, results_dev, st))) return rc;
    return 0;
}

int vfsms_align_batch_dev(vfsms_ctx *ctx, const uint8_t *rois_a_dev, const uint8_t *rois_b_dev, int n_pairs,
                          int rows, int cols, int stride, int64_t pair_stride, const vfsms_surf_params *params,
                          float ratio, int offset_evaluate, vfsms_pair_result *results_dev, void *stream)
{
    if (!ctx || !rois_a_dev || !rois_b_dev || !params || !results_dev || n_pairs < 1) { vfsms_set_error("align_batch_dev: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    int rc;
    if ((rc = surf_reserve(ctx, 2 * n_pairs, rows, cols, params))) return rc;
    if ((rc = match_reserve(ctx, n_pairs, ctx->surf.kp_cap))) return rc;
    return align_batch_launch(ctx, rois_a_dev, rois_b_dev, n_pairs, rows, cols, stride, pair_stride, params, (double)ratio,
                              offset_evaluate, results_dev, st);
}

// detect + describe + match + vote on device-resident ROIs, results to the host; regrows the candidate / keypoint buffers
// and repeats when a pair reports overflow
static int align_dev_regrow(vfsms_ctx *ctx, const uint8_t *a_dev, const uint8_t *b_dev, int n_pairs, int rows, int cols, int stride,
                            int64_t pair_stride, const vfsms_surf_params *params, float ratio, int offset_evaluate,
                            vfsms_pair_result *results, cudaStream_t st)
{
    int rc;
    if ((rc = ctx->results.reserve(sizeof(vfsms_pair_result) * (size_t)n_pairs))) return rc;
    if ((rc = surf_reserve(ctx, 2 * n_pairs, rows, cols, params))) return rc;
    for (int attempt = 0; attempt < 6; attempt++) {
        if ((rc = match_reserve(ctx, n_pairs, ctx->surf.kp_cap))) return rc;
        if ((rc = align_batch_launch(ctx, a_dev, b_dev, n_pairs, rows, cols, stride, pair_stride,
                                     params, (double)ratio, offset_evaluate, ctx->results.as<vfsms_pair_result>(), st))) return rc;
        CUDA_TRY(cudaMemcpyAsync(results, ctx->results.p, sizeof(vfsms_pair_result) * (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        bool overflow = false;
        for (int p = 0; p < n_pairs; p++) overflow |= (results[p].flags & 1) != 0;
        if (!overflow) return 0;
        std::vector<int32_t> cnt((size_t)8 * n_pairs);
        CUDA_TRY(cudaMemcpy(cnt.data(), ctx->surf.counters.p, cnt.size() * 4, cudaMemcpyDeviceToHost));
        int gc = 0, gk = 0;
        for (int b = 0; b < 2 * n_pairs; b++) { gc |= cnt[b * 4 + 3] & 1; gk |= cnt[b * 4 + 3] & 2; }
        if ((rc = surf_grow(ctx, gc, gk))) return rc;
    }
    vfsms_set_error("align: candidate buffer overflow after regrow");
    return VFSMS_E_OVERFLOW;
}

int vfsms_align_batch_host(vfsms_ctx *ctx, const uint8_t *rois_a, const uint8_t *rois_b, int n_pairs,
                           int rows, int cols, int stride, int64_t pair_stride, const vfsms_surf_params *params,
                           float ratio, int offset_evaluate, vfsms_pair_result *results)
{
    if (!ctx || !rois_a || !rois_b || !params || !results || n_pairs < 1 || rows < 1 || cols < 1) { vfsms_set_error("align_batch_host: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    const size_t img_bytes = (size_t)rows * cols;
    if ((rc = ctx->img_a.reserve(img_bytes * n_pairs))) return rc;
    if ((rc = ctx->img_b.reserve(img_bytes * n_pairs))) return rc;
    if (stride == cols && pair_stride == (int64_t)img_bytes) {          // packed ROIs: one copy per side
        CUDA_TRY(cudaMemcpyAsync(ctx->img_a.p, rois_a, img_bytes * n_pairs, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ctx->img_b.p, rois_b, img_bytes * n_pairs, cudaMemcpyHostToDevice, st));
    } else
    for (int p = 0; p < n_pairs; p++) {
        CUDA_TRY(cudaMemcpy2DAsync(ctx->img_a.as<uint8_t>() + p * img_bytes, cols, rois_a + p * pair_stride, stride, cols, rows, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpy2DAsync(ctx->img_b.as<uint8_t>() + p * img_bytes, cols, rois_b + p * pair_stride, stride, cols, rows, cudaMemcpyHostToDevice, st));
    }
    return align_dev_regrow(ctx, ctx->img_a.as<uint8_t>(), ctx->img_b.as<uint8_t>(), n_pairs, rows, cols, cols, (int64_t)img_bytes, params, ratio,
                            offset_evaluate, results, st);
}

int vfsms_align_batch_upload(vfsms_ctx *ctx, int slot, const uint8_t *rois_a, const uint8_t *rois_b, int n_pairs,
                             int rows, int cols, int stride, int64_t pair_stride)
{
    if (!ctx || slot < 0 || slot > 1 || !rois_a || !rois_b || n_pairs < 1 || rows < 1 || cols < 1) { vfsms_set_error("align_batch_upload: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    vfsms_ctx::UploadSlot &s = ctx->slots[slot];
    if (!s.uploaded) CUDA_TRY(cudaEventCreateWithFlags(&s.uploaded, cudaEventDisableTiming));
    int rc;
    const size_t img_bytes = (size_t)rows * cols;
    // a slot is written again only after vfsms_align_batch_run of its previous batch returned (that call synchronises)
    if ((rc = s.a.reserve(img_bytes * n_pairs))) return rc;
    if ((rc = s.b.reserve(img_bytes * n_pairs))) return rc;
    if (stride == cols && pair_stride == (int64_t)img_bytes) {
        CUDA_TRY(cudaMemcpyAsync(s.a.p, rois_a, img_bytes * n_pairs, cudaMemcpyHostToDevice, ctx->copy_stream));
        CUDA_TRY(cudaMemcpyAsync(s.b.p, rois_b, img_bytes * n_pairs, cudaMemcpyHostToDevice, ctx->copy_stream));
    } else {
        for (int p = 0; p < n_pairs; p++) {
            CUDA_TRY(cudaMemcpy2DAsync(s.a.as<uint8_t>() + p * img_bytes, cols, rois_a + p * pair_stride, stride, cols, rows, cudaMemcpyHostToDevice, ctx->copy_stream));
            CUDA_TRY(cudaMemcpy2DAsync(s.b.as<uint8_t>() + p * img_bytes, cols, rois_b + p * pair_stride, stride, cols, rows, cudaMemcpyHostToDevice, ctx->copy_stream));
        }
    }
    CUDA_TRY(cudaEventRecord(s.uploaded, ctx->copy_stream));
    s.n_pairs = n_pairs; s.rows = rows; s.cols = cols;
    return 0;
}

int vfsms_align_batch_run(vfsms_ctx *ctx, int slot, const vfsms_surf_params *params, float ratio, int offset_evaluate,
                          vfsms_pair_result *results)
{
    if (!ctx || slot < 0 || slot > 1 || !params || !results) { vfsms_set_error("align_batch_run: bad arguments"); return VFSMS_E_ARG; }
    vfsms_ctx::UploadSlot &s = ctx->slots[slot];
    if (!s.uploaded || s.n_pairs < 1) { vfsms_set_error("align_batch_run: slot %d holds no uploaded batch", slot); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamWaitEvent(ctx->stream, s.uploaded, 0));
    const int n_pairs = s.n_pairs;
    s.n_pairs = 0;                      // consumed: a second run needs a new upload
    return align_dev_regrow(ctx, s.a.as<uint8_t>(), s.b.as<uint8_t>(), n_pairs, s.rows, s.cols, s.cols, (int64_t)s.rows * s.cols, params, ratio,
                            offset_evaluate, results, ctx->stream);
}

/* ---------------------------------------------------------------- device-resident tile stack */
int vfsms_tiles_reserve(vfsms_ctx *ctx, int n_tiles, int rows, int cols)
{
    if (!ctx || n_tiles < 1 || rows < 1 || cols < 1) { vfsms_set_error("tiles_reserve: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc;
    if (ctx->tiles_borrowed) { ctx->tiles.p = nullptr; ctx->tiles.bytes = 0; ctx->tiles_borrowed = false; }     // drop the alias, own a stack again
    if ((rc = ctx->tiles.reserve((size_t)n_tiles * rows * cols))) return rc;
    ctx->tiles_n = n_tiles; ctx->tiles_rows = rows; ctx->tiles_cols = cols;
    ctx->tiles_has_bgr.assign((size_t)n_tiles, 0);   // a new reservation starts without colour tiles
    return 0;
}

int vfsms_tiles_attach(vfsms_ctx *ctx, vfsms_ctx *owner)
{
    if (!ctx || !owner || ctx == owner || !owner->tiles.p || ctx->device != owner->device) { vfsms_set_error("tiles_attach: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(owner->stream));          // uploads / decodes into the stack have landed
    if (!ctx->tiles_borrowed) ctx->tiles.release();
    ctx->tiles.p = owner->tiles.p; ctx->tiles.bytes = owner->tiles.bytes; ctx->tiles_borrowed = true;
    ctx->tiles_n = owner->tiles_n; ctx->tiles_rows = owner->tiles_rows; ctx->tiles_cols = owner->tiles_cols;
    ctx->tiles_has_bgr.assign((size_t)ctx->tiles_n, 0);
    return 0;
}

static int tiles_range_ok(vfsms_ctx *ctx, int first, int n, const char *who)
{
    if (!ctx || !ctx->tiles.p || first < 0 || n < 0 || first + n > ctx->tiles_n) {
        vfsms_set_error("%s: tiles [%d, %d) outside the reserved stack of %d", who, first, first + n, ctx ? ctx->tiles_n : 0); return VFSMS_E_ARG;
    }
    return 0;
}

int vfsms_tiles_decode_jpeg(vfsms_ctx *ctx, int first, int n, const uint8_t *const *data, const size_t *sizes)
{
    int rc;
    if ((rc = tiles_range_ok(ctx, first, n, "tiles_decode_jpeg"))) return rc;
    const int64_t img = (int64_t)ctx->tiles_rows * ctx->tiles_cols;
    return vfsms_jpeg_decode_gray_dev(ctx, n, data, sizes, ctx->tiles.as<uint8_t>() + first * img, ctx->tiles_rows, ctx->tiles_cols,
                                      ctx->tiles_cols, img, nullptr);
}

static int tiles_bgr_reserve(vfsms_ctx *ctx)
{
    const size_t need = (size_t)ctx->tiles_n * ctx->tiles_rows * ctx->tiles_cols * 3;
    if (ctx->tiles_bgr.bytes >= need) return 0;
    // growing would move the colour tiles already present: the twin is sized for the whole reservation at once
    std::fill(ctx->tiles_has_bgr.begin(), ctx->tiles_has_bgr.end(), 0);
    return ctx->tiles_bgr.reserve(need);
}

int vfsms_tiles_decode_jpeg_bgr(vfsms_ctx *ctx, int first, int n, const uint8_t *const *data, const size_t *sizes)
{
    int rc;
    if ((rc = tiles_range_ok(ctx, first, n, "tiles_decode_jpeg_bgr"))) return rc;
    if (!data || !sizes) { vfsms_set_error("tiles_decode_jpeg_bgr: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    if ((rc = tiles_bgr_reserve(ctx))) return rc;
    const int64_t img = (int64_t)ctx->tiles_rows * ctx->tiles_cols;
    rc = jpeg_decode_bgr_gray_dev(ctx, n, data, sizes, ctx->tiles_bgr.as<uint8_t>() + first * img * 3, ctx->tiles_rows, ctx->tiles_cols,
                                  (int64_t)ctx->tiles_cols * 3, img * 3, ctx->tiles.as<uint8_t>() + first * img, ctx->tiles_cols, img, ctx->stream);
    if (rc) return rc;
    for (int k = first; k < first + n; k++) ctx->tiles_has_bgr[k] = 1;
    return 0;
}

int vfsms_tiles_upload_bgr(vfsms_ctx *ctx, int first, int n, const uint8_t *tiles_bgr)
{
    int rc;
    if ((rc = tiles_range_ok(ctx, first, n, "tiles_upload_bgr"))) return rc;
    if (!tiles_bgr) { vfsms_set_error("tiles_upload_bgr: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    if ((rc = tiles_bgr_reserve(ctx))) return rc;
    const size_t img = (size_t)ctx->tiles_rows * ctx->tiles_cols * 3;
    CUDA_TRY(cudaMemcpyAsync(ctx->tiles_bgr.as<uint8_t>() + first * img, tiles_bgr, img * n, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int k = first; k < first + n; k++) ctx->tiles_has_bgr[k] = 1;
    return 0;
}

int vfsms_tiles_upload(vfsms_ctx *ctx, int first, int n, const uint8_t *tiles)
{
    int rc;
    if ((rc = tiles_range_ok(ctx, first, n, "tiles_upload"))) return rc;
    if (!tiles) { vfsms_set_error("tiles_upload: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t img = (size_t)ctx->tiles_rows * ctx->tiles_cols;
    CUDA_TRY(cudaMemcpyAsync(ctx->tiles.as<uint8_t>() + first * img, tiles, img * n, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int vfsms_tiles_download(vfsms_ctx *ctx, int first, int n, uint8_t *out)
{
    int rc;
    if ((rc = tiles_range_ok(ctx, first, n, "tiles_download"))) return rc;
    if (!out) { vfsms_set_error("tiles_download: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t img = (size_t)ctx->tiles_rows * ctx->tiles_cols;
    CUDA_TRY(cudaMemcpyAsync(out, ctx->tiles.as<uint8_t>() + first * img, img * n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

const uint8_t *vfsms_tiles_ptr(vfsms_ctx *ctx) { return ctx ? ctx->tiles.as<uint8_t>() : nullptr; }

int vfsms_tiles_align_strided(vfsms_ctx *ctx, int first, int n_pairs, int pair_step, int direction, int roi_len,
                              const vfsms_surf_params *params, float ratio, int offset_evaluate, vfsms_pair_result *results)
{
    int rc;
    if (pair_step < 1 || n_pairs < 1) { vfsms_set_error("tiles_align: bad arguments"); return VFSMS_E_ARG; }
    if ((rc = tiles_range_ok(ctx, first, (n_pairs - 1) * pair_step + 2, "tiles_align"))) return rc;
    const int H = ctx->tiles_rows, W = ctx->tiles_cols;
    const int edge = (direction == 1 || direction == 3) ? H : W;
    if (!params || !results || direction < 1 || direction > 4 || roi_len < 1 || roi_len > edge) {
        vfsms_set_error("tiles_align: bad arguments"); return VFSMS_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    // ROI of the first image / of the second image (ImageUtility.py:66-101), read in place from the stack
    const int64_t img = (int64_t)H * W;
    const uint8_t *A = ctx->tiles.as<uint8_t>() + first * img, *B = A + img;
    int rows, cols;
    if (direction == 1) { rows = roi_len; cols = W; A += (int64_t)(H - roi_len) * W; }            // A bottom strip, B top strip
    else if (direction == 2) { rows = H; cols = roi_len; A += W - roi_len; }                      // A right strip, B left strip
    else if (direction == 3) { rows = roi_len; cols = W; B += (int64_t)(H - roi_len) * W; }       // A top strip, B bottom strip
    else { rows = H; cols = roi_len; B += W - roi_len; }                                          // A left strip, B right strip
    return align_dev_regrow(ctx, A, B, n_pairs, rows, cols, W, img * pair_step, params, ratio, offset_evaluate, results, ctx->stream);
}

int vfsms_tiles_align_list(vfsms_ctx *ctx, int n_pairs, const int32_t *first_tile, const int32_t *direction, int roi_len,
                           const vfsms_surf_params *params, float ratio, int offset_evaluate, vfsms_pair_result *results)
{
    if (!ctx || !first_tile || !direction || !params || !results || n_pairs < 1) { vfsms_set_error("tiles_align_list: bad arguments"); return VFSMS_E_ARG; }
    const int H = ctx->tiles_rows, W = ctx->tiles_cols;
    const bool tall = direction[0] == 2 || direction[0] == 4;            // column strips (H x roi_len) vs row strips (roi_len x W)
    const int edge = tall ? W : H;
    if (roi_len < 1 || roi_len > edge) { vfsms_set_error("tiles_align_list: bad ROI length"); return VFSMS_E_ARG; }
    const int64_t img = (int64_t)H * W;
    std::vector<int64_t> off((size_t)2 * n_pairs);
    int rc;
    for (int p = 0; p < n_pairs; p++) {
        const int d = direction[p];
        if (d < 1 || d > 4 || ((d == 2 || d == 4) != tall)) { vfsms_set_error("tiles_align_list: directions of one call must share the strip shape"); return VFSMS_E_ARG; }
        if ((rc = tiles_range_ok(ctx, first_tile[p], 2, "tiles_align_list"))) return rc;
        int64_t a = first_tile[p] * img, b = a + img;                    // ImageUtility.py:66-101
        if (d == 1) a += (int64_t)(H - roi_len) * W;
        else if (d == 2) a += W - roi_len;
        else if (d == 3) b += (int64_t)(H - roi_len) * W;
        else b += W - roi_len;
        off[p] = a; off[(size_t)n_pairs + p] = b;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    if ((rc = ctx->surf.img_off_buf.reserve(off.size() * 8))) return rc;
    CUDA_TRY(cudaMemcpyAsync(ctx->surf.img_off_buf.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));                         // `off` is pageable host memory
    ctx->surf.img_off = ctx->surf.img_off_buf.as<int64_t>();
    const uint8_t *base = ctx->tiles.as<uint8_t>();
    rc = align_dev_regrow(ctx, base, base, n_pairs, tall ? H : roi_len, tall ? roi_len : W, W, img, params, ratio, offset_evaluate, results, ctx->stream);
    ctx->surf.img_off = nullptr;
    return rc;
}

int vfsms_tiles_align(vfsms_ctx *ctx, int first, int n_pairs, int direction, int roi_len, const vfsms_surf_params *params, float ratio,
                      int offset_evaluate, vfsms_pair_result *results)
{
    return vfsms_tiles_align_strided(ctx, first, n_pairs, 1, direction, roi_len, params, ratio, offset_evaluate, results);
}

int vfsms_match_batch_dev(vfsms_ctx *ctx, const float *desc_a_dev, const int32_t *n_a_dev, const float *desc_b_dev,
                          const int32_t *n_b_dev, int n_pairs, int cap, int dim, float ratio,
                          int32_t *best_idx_dev, float *best_dist_dev, void *stream)
{
    (void)ratio;
    if (!ctx || cap % 64) { vfsms_set_error("match_batch_dev: cap must be a multiple of 64"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    int rc;
    if ((rc = ctx->scratch2.reserve((size_t)2 * n_pairs * cap * dim * 4))) return rc;
    float *AT = ctx->scratch2.as<float>(), *BT = AT + (size_t)n_pairs * cap * dim;
    CUDA_TRY(cudaMemsetAsync(AT, 0, (size_t)2 * n_pairs * cap * dim * 4, st));
    if ((rc = transpose_desc_batch(ctx, desc_a_dev, n_a_dev, 1, AT, n_pairs, cap, dim, (int64_t)cap * dim, (int64_t)cap * dim, st))) return rc;
    if ((rc = transpose_desc_batch(ctx, desc_b_dev, n_b_dev, 1, BT, n_pairs, cap, dim, (int64_t)cap * dim, (int64_t)cap * dim, st))) return rc;
    return match_l2_knn2_batch(ctx, AT, n_a_dev, 1, BT, n_b_dev, 1, n_pairs, cap, dim, (int64_t)cap * dim, (int64_t)cap * dim,
                               best_idx_dev, best_dist_dev, st);
}

}  // extern "C"
