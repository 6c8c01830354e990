/*
 * vfsms.h -- C ABI of libvfsms.so: the B200-native replacement for the pairwise-alignment hot path of
 * Keep-Passion/ImageStitch (VFSMS).  Plain pointers and sizes only; no torch / numpy / OpenCV types.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the reference tree).
 * The reference's own native boundary is the Boost.Python module `myGpuFeatures`
 * (appendix/myGpuFeatures.cpp:203-209) with three functions; the first block below is that boundary,
 * the second block is the fused / batched path the Python host (imagestitch_b200/) drives.
 *
 * Conventions
 *   - all functions return 0 on success or a negative VFSMS_E_* code; vfsms_last_error() gives the text
 *     (thread-local).  Nothing here falls back to a CPU implementation: without a CUDA device
 *     vfsms_create() fails with VFSMS_E_NODEVICE.
 *   - "host" pointers are ordinary (preferably pinned) host memory; "dev" pointers are device memory on the
 *     context's device.  `stream` is a cudaStream_t passed as void* (NULL = the context's own stream).
 *   - keypoint records are 8 float32: x, y, size, angle, response, octave, laplacian, 0   (VFSMS_KP_STRIDE).
 *   - row = first image axis (the reference calls it "dx"), col = second axis ("dy"), ImageUtility.py:154-161.
 */
#ifndef VFSMS_H
#define VFSMS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFSMS_VERSION 100
#define VFSMS_KP_STRIDE 8

enum {
    VFSMS_OK = 0,
    VFSMS_E_NODEVICE = -1,   /* no CUDA device / driver */
    VFSMS_E_CUDA = -2,       /* CUDA runtime error (text in vfsms_last_error) */
    VFSMS_E_ARG = -3,        /* invalid argument */
    VFSMS_E_CAPACITY = -4,   /* caller-provided output capacity too small (needed size reported) */
    VFSMS_E_OVERFLOW = -5,   /* internal candidate buffer overflow even after regrow */
    VFSMS_E_UNSUPPORTED = -6
};

typedef struct vfsms_ctx vfsms_ctx;

/* SURF parameters: the argument list of myGpuFeatures.detectAndDescribeBySurf (appendix/myGpuFeatures.cpp:67,
 * defaults ImageUtility.py:23-28) / cv2.xfeatures2d.SURF_create (ImageUtility.py:258). */
typedef struct {
    float hessian_threshold;   /* 100 */
    int n_octaves;             /* 4 */
    int n_octave_layers;       /* 3 */
    int extended;              /* 1 -> 128-d, 0 -> 64-d */
    float keypoints_ratio;     /* GPU semantics: keep at most min(ratio*rows*cols, 65535) strongest; <= 0: unlimited */
    int upright;               /* 0 */
} vfsms_surf_params;

/* Per-pair result of the fused alignment (one evaluation of the body of the search loop,
 * Stitcher.py:319-346: detect x2 -> match -> getOffsetByMode). */
typedef struct {
    int32_t status;     /* 1 when votes >= offset_evaluate (ImageUtility.py:175) */
    int32_t d_row;      /* ROI-relative mode offset, before the ROI origin is added back (Stitcher.py:353-360) */
    int32_t d_col;
    int32_t votes;
    int32_t n_a;        /* keypoints in ROI A */
    int32_t n_b;        /* keypoints in ROI B */
    int32_t n_matches;  /* ratio-test survivors */
    int32_t flags;      /* bit0: candidate overflow (result invalid) */
} vfsms_pair_result;

/* ---------------------------------------------------------------- library / context */
int vfsms_version(void);
const char *vfsms_last_error(void);
int vfsms_device_count(void);                 /* 0 when no usable CUDA device */
/* One context per (process, device): owns a stream, workspaces and cuFFT plans. */
int vfsms_create(int device, vfsms_ctx **out);
void vfsms_destroy(vfsms_ctx *ctx);
int vfsms_synchronize(vfsms_ctx *ctx);
void *vfsms_stream(vfsms_ctx *ctx);           /* the context's cudaStream_t */
/* number of kernel launches this context has issued since creation (bench.py's gpu_launches) */
int64_t vfsms_launch_count(vfsms_ctx *ctx);

/* Matcher selection.  All modes produce identical results (candidates from a tcgen05 GEMM on reduced-precision operands,
 * exact fp32 rescoring, and an exact rescan of every query whose candidates the operand error bound cannot separate):
 *   0 (default) = CTA pairs (cta_group::2, 256 x 256 tiles) with the default operand scheme (plain fp16, mode 5),
 *   1 = exact fp32 SIMT kernel, no tensor cores (verification),
 *   2 = like 0 on single CTAs (128 x 128 tiles, cta_group::1) (verification),
 *   3 / 4 / 5 = like 0 with split-bf16 operands (3 product terms, K' = 3D + 16) / fp16 with the query split
 *               (2 terms, K' = 2D + 16) / plain fp16 (1 term, K' = D + 16).
 * The fp16 schemes need squared descriptor norms <= 1e4 (SURF descriptors are unit vectors); a pair that breaks this is
 * rescanned exactly, and vfsms_match uses the bf16 scheme for featureType 1 (SIFT) descriptors. */
int vfsms_set_matcher(vfsms_ctx *ctx, int mode);
/* Kernel-schedule switches.  Every value of an option produces identical results; the defaults are what bench.py times and what
 * the oracle tests run on, the others stay selectable for verification and A/B measurement (bench.py --opt name=value, environment
 * VFSMS_OPTS="name=value,...": read once per vfsms_create).  Unknown options / values return VFSMS_E_ARG. */
enum {
    VFSMS_OPT_DESCRIBE_MODE = 0,  /* "describe": rotated-window sampler of the SURF descriptor (csrc/surf_describe.cuh).
                                   * 1 (default) = 32.32 fixed-point positions, window rows sampled in chunks through one stacked float
                                   * texture with four gathers in flight per lane; keypoints whose direction or row starts are not
                                   * multiples of 2^-32 are handed to sampler 0.  0 = reference sampler for every keypoint: one row at
                                   * a time, double precision, straight from the u8 image.  Values 0 and 1 give bit-identical results.
                                   * 2 and 3 are TOLERANCE modes (keypoints identical, descriptors not): 2 = float positions, fp32
                                   * lerps on the gathered footprint (95 % of the descriptors identical, max |diff| 0.06, mean 1e-6;
                                   * -5 % describe time); 3 = the texture unit's own bilinear filter with its 8-bit weights (max |diff|
                                   * 0.15, mean 2e-4, >= 97 % of the matches identical, offsets identical on every fixture; -12 %).
                                   * tests/test_gpu_variants.py asserts these tolerances; DESIGN.md 12 */
    VFSMS_OPT_SORT_MODE = 1,      /* "sort": KeypointGreater ordering.  1 (default) = 13-bit response histogram, then rank counting
                                   * inside each candidate's own bin; 0 = rank by counting over all staged candidates */
    VFSMS_OPT_DESCRIBE_LPT = 2,   /* "lpt": describe the large windows first (two passes over the work list), so that no giant window
                                   * is met at the end of the launch: split at 128 px (1, default), 64 px (2), 256 px (3); 0 = keypoints
                                   * in response order */
    VFSMS_OPT_ENTROPY = 3,        /* "entropy": Huffman decoding of JPEG tiles.  1 (default) = on the device: self-synchronising
                                   * parallel decode of 1024-bit subsequences (files with restart intervals take the host stage);
                                   * 0 = host threads, one file each; same coefficients, symbol for symbol */
    VFSMS_OPT_COUNT
};
int vfsms_set_option(vfsms_ctx *ctx, int option, int value);
int vfsms_get_option(vfsms_ctx *ctx, int option, int *value_out);
const char *vfsms_option_name(int option);
/* Number of queries of the last tensor-core match that needed the exact fallback scan (synchronises the stream). */
int vfsms_last_match_fallbacks(vfsms_ctx *ctx, int *count_out);
/* Number of rescored candidates of the last tensor-core match whose GEMM score was further from the exact score than the
 * error bound the guard relies on; 0 unless the bound is wrong (synchronises the stream). */
int vfsms_last_match_bound_violations(vfsms_ctx *ctx, int *count_out);
/* Number of keypoints of the last SURF run that the fixed-point window sampler handed to the reference sampler (synchronises). */
int vfsms_last_describe_handovers(vfsms_ctx *ctx, int *count_out);

/* Per-stage device timing with CUDA events recorded on the launching stream (bench.py's roofline numbers).
 * Off by default.  vfsms_profile_read synchronises the stream, then returns accumulated milliseconds and call
 * counts per stage (arrays of VFSMS_STAGE_COUNT) and optionally resets them. */
enum {
    VFSMS_STAGE_INTEGRAL = 0, VFSMS_STAGE_HESSIAN, VFSMS_STAGE_SORT, VFSMS_STAGE_COMPACT, VFSMS_STAGE_DESCRIBE,
    VFSMS_STAGE_TRANSPOSE, VFSMS_STAGE_MATCH, VFSMS_STAGE_VOTE, VFSMS_STAGE_PHASE_FFT, VFSMS_STAGE_PHASE_PEAK,
    VFSMS_STAGE_BLEND, VFSMS_STAGE_MATCH_TC, VFSMS_STAGE_COUNT
};
int vfsms_profile_enable(vfsms_ctx *ctx, int on);
int vfsms_profile_read(vfsms_ctx *ctx, float *ms_out, int32_t *calls_out, int reset);
const char *vfsms_stage_name(int stage);

/* ---------------------------------------------------------------- legacy plugin boundary (host in, host out) */

/* Replaces myGpuFeatures.detectAndDescribeBySurf (appendix/myGpuFeatures.cpp:67-104; caller ImageUtility.py:272).
 * image: rows x cols u8, row stride `stride` bytes (strided ROI views are accepted like NDArrayConverter::toMat,
 * appendix/conversion.cpp:145-243).  kp_out: cap x 8 float32, desc_out: cap x (64|128) float32.
 * *n_out = keypoints found; VFSMS_E_CAPACITY (and *n_out = needed) when cap is too small. */
int vfsms_surf_detect_and_describe(vfsms_ctx *ctx, const uint8_t *image, int rows, int cols, int stride,
                                   const vfsms_surf_params *params, float *kp_out, float *desc_out, int cap,
                                   int *n_out);

/* Replaces myGpuFeatures.matchDescriptors (appendix/myGpuFeatures.cpp:148-195; callers ImageUtility.py:306,308).
 * feature_type 1|2: L2 kNN(2) + `d0 < param * d1` (cpp:160-173); 3: Hamming best-1 + `d < param` (cpp:174-187,
 * descriptors are byte values stored as float32, cpp:118).  matches_out: n_a x 2 int32 rows (trainIdx, queryIdx)
 * in ascending queryIdx (cpp:53-65). */
int vfsms_match_descriptors(vfsms_ctx *ctx, const float *desc_a, int n_a, const float *desc_b, int n_b, int dim,
                            int feature_type, float param, int32_t *matches_out, int *m_out);

/* Replaces myGpuFeatures.detectAndDescribeByOrb (appendix/myGpuFeatures.cpp:106-146; caller ImageUtility.py:274).
 * desc_out: cap x 32 float32 holding byte values 0..255 (cpp:118). */
int vfsms_orb_detect_and_describe(vfsms_ctx *ctx, const uint8_t *image, int rows, int cols, int stride,
                                  int n_features, float scale_factor, int n_levels, int edge_threshold,
                                  int first_level, int wta_k, int patch_size, int fast_threshold,
                                  float *kp_out, float *desc_out, int cap, int *n_out);

/* ---------------------------------------------------------------- fused alignment path */

/* Replaces Method.getOffsetByMode (ImageUtility.py:139-178).  kps_*: n x kp_stride float32 with (x, y) first;
 * matches: m x 2 int32 (trainIdx, queryIdx).  result: status, d_row, d_col, votes filled. */
int vfsms_offset_by_mode(vfsms_ctx *ctx, const float *kps_a, int n_a, const float *kps_b, int n_b, int kp_stride,
                         const int32_t *matches, int m, int offset_evaluate, vfsms_pair_result *result);

/* One body of the incremental search loop for a batch of ROI pairs (Stitcher.py:323-346 with
 * featureMethod="surf", offsetCaculate="mode"):  results[p] = vote(match(surf(roi_a[p]), surf(roi_b[p]))).
 * All ROIs of a batch share rows x cols; roi_*[p] starts at base + p * pair_stride bytes, row stride `stride`.
 * Host variant: H2D of the ROIs and D2H of the results are inside the call (what the reference boundary includes,
 * appendix/myGpuFeatures.cpp:74,84).  Dev variant: inputs already in HBM, results written to device memory,
 * asynchronous on `stream`. */
int vfsms_align_batch_host(vfsms_ctx *ctx, const uint8_t *rois_a, const uint8_t *rois_b, int n_pairs,
                           int rows, int cols, int stride, int64_t pair_stride,
                           const vfsms_surf_params *params, float ratio, int offset_evaluate,
                           vfsms_pair_result *results);
int vfsms_align_batch_dev(vfsms_ctx *ctx, const uint8_t *rois_a_dev, const uint8_t *rois_b_dev, int n_pairs,
                          int rows, int cols, int stride, int64_t pair_stride,
                          const vfsms_surf_params *params, float ratio, int offset_evaluate,
                          vfsms_pair_result *results_dev, void *stream);
/* Double-buffered form of vfsms_align_batch_host for a stream of batches.  _upload enqueues the host -> device copy of one
 * batch into input slot 0 or 1 on the context's copy stream and returns at once (pinned host memory makes the copy
 * asynchronous; the buffers must stay valid until the matching _run returned).  _run waits for that slot's copy, aligns the
 * batch exactly like vfsms_align_batch_host and returns the results.  Uploading batch k+1 into the other slot before running
 * batch k hides its copy behind batch k's kernels.  A slot may be uploaded again once its _run has returned. */
int vfsms_align_batch_upload(vfsms_ctx *ctx, int slot, const uint8_t *rois_a, const uint8_t *rois_b, int n_pairs,
                             int rows, int cols, int stride, int64_t pair_stride);
int vfsms_align_batch_run(vfsms_ctx *ctx, int slot, const vfsms_surf_params *params, float ratio, int offset_evaluate,
                          vfsms_pair_result *results);

/* Matcher on device-resident descriptors (bench / profiling hook for the BF matcher, ImageUtility.py:278-309):
 * exact fp32 kNN(2)+ratio for n_pairs descriptor sets laid out [pair][cap][dim], counts per pair on device. */
int vfsms_match_batch_dev(vfsms_ctx *ctx, const float *desc_a_dev, const int32_t *n_a_dev, const float *desc_b_dev,
                          const int32_t *n_b_dev, int n_pairs, int cap, int dim, float ratio,
                          int32_t *best_idx_dev /* [pair][cap][2] */, float *best_dist_dev /* [pair][cap][2] */,
                          void *stream);

/* Optional ROI enhancement before detection (Stitcher.py:269-276, 327-334; off by default).
 * mode 0: cv2.equalizeHist(image);  mode 1: cv2.createCLAHE(clip_limit, (tile_grid, tile_grid)).apply(image)
 * (reference values: clipLimit 20, tileSize 5, ImageUtility.py:47-50).  out: rows x cols u8, contiguous. */
int vfsms_enhance_host(vfsms_ctx *ctx, const uint8_t *image, int rows, int cols, int stride, int mode, double clip_limit,
                       int tile_grid, uint8_t *out);

/* ---------------------------------------------------------------- phase correlation */

/* Replaces cv2.phaseCorrelate(np.float64(roiA), np.float64(roiB)) as called at Stitcher.py:230 (no window):
 * pads to the optimal DFT size, cross-power spectrum, inverse transform, 5x5 weighted centroid.
 * out: shift_x, shift_y, response (float64, cv2's sign convention). */
int vfsms_phase_correlate_host(vfsms_ctx *ctx, const uint8_t *roi_a, const uint8_t *roi_b, int rows, int cols,
                               int stride, double out[3]);
int vfsms_phase_correlate_dev(vfsms_ctx *ctx, const uint8_t *roi_a_dev, const uint8_t *roi_b_dev, int rows, int cols,
                              int stride, double *out_dev, void *stream);

/* Scoring step of the opt-in wrap-aware phase mode (imagestitch_b200/phase_wrap.py -- the reference has no
 * such check, its phase path Stitcher.py:205-258 adds the shift with the wrong sign and ignores aliasing, quirk Q6).
 * For each candidate shift (dRow, dCol) with roiB(r, c) <-> roiA(r + dRow, c + dCol): the integer sums n, Sa, Sb, Sab, Saa, Sbb over
 * the pixels both ROIs share.  shifts: n_shifts x 2 int32 (n_shifts <= 64), sums_out: n_shifts x 6 int64.  Exact integers. */
int vfsms_overlap_sums_host(vfsms_ctx *ctx, const uint8_t *roi_a, const uint8_t *roi_b, int rows, int cols, int stride_a, int stride_b,
                            int n_shifts, const int32_t *shifts, int64_t *sums_out);

/* ---------------------------------------------------------------- overlap blending */

enum {
    VFSMS_FUSE_NONE = 0,        /* "notFuse"            Stitcher.py:507 */
    VFSMS_FUSE_AVERAGE = 1,     /* fuseByAverage        ImageFusion.py:12-21 */
    VFSMS_FUSE_MAXIMUM = 2,     /* fuseByMaximum        ImageFusion.py:23-31 */
    VFSMS_FUSE_MINIMUM = 3,     /* fuseByMinimum        ImageFusion.py:33-41 */
    VFSMS_FUSE_FADE = 4,        /* fuseByFadeInAndFadeOut ImageFusion.py:192-244 */
    VFSMS_FUSE_TRIG = 5,        /* fuseByTrigonometric  ImageFusion.py:246-293 */
    VFSMS_FUSE_MULTIBAND = 6    /* fuseByMultiBandBlending ImageFusion.py:296-367 */
};

/* Replaces Stitcher.fuseImage (Stitcher.py:488-525) for one overlap ROI.  a, b: rows x cols x channels int16 with
 * -1 = empty (the reference's int64 canvas sentinel, Stitcher.py:434-436); out: rows x cols x channels u8.
 * d_row / d_col: the ORIGINAL pair offset whose sign selects the ramp direction (Stitcher.py:478,483).
 * method | 0x100 forces the corner-weight path of getWeightsMatrix regardless of the fill ratio (parity hook). */
int vfsms_fuse_roi_host(vfsms_ctx *ctx, const int16_t *a, const int16_t *b, int rows, int cols, int channels,
                        int method, int d_row, int d_col, uint8_t *out,
                        float *weight_a_out /* rows x cols float32 or NULL: ImageFusion.getWeightsMatrix parity hook */,
                        float *weight_b_out);

/* Replaces Stitcher.getStitchByOffset's paste/blend loop (Stitcher.py:433-486) on a device-resident canvas.
 * tiles: n_tiles images of tile_rows x tile_cols x channels u8 (host), placed at rectified offsets (row, col);
 * occupied-bbox bookkeeping (rangeX / rangeY) is computed by the host mirror and passed as roi rects.
 * canvas_out: canvas_rows x canvas_cols x channels u8 (host). */
int vfsms_mosaic_host(vfsms_ctx *ctx, const uint8_t *tiles, int n_tiles, int tile_rows, int tile_cols, int channels,
                      const int32_t *tile_origin /* n x 2 */, const int32_t *roi_rect /* n x 4: r0,c0,r1,c1 */,
                      const int32_t *pair_offset /* n x 2 original offsets */, int method,
                      int canvas_rows, int canvas_cols, uint8_t *canvas_out);

/* One BAND of a mosaic whose tile sequence is partitioned over GPUs (the reference's loop Stitcher.py:440-483 is
 * strictly sequential).  A band owns a contiguous run of tiles and a canvas = the bounding box of their rectangles; all
 * coordinates are band-local.  halo_in (int16, -1 = empty, halo_in_rect = r0, c0, rows, cols) is what the previous bands left
 * inside this canvas; it is pasted before the first tile so that every ROI sees exactly the pixels the sequential loop would
 * (the corner-weight scans of ImageFusion.py:62-187 depend on them).  fuse_first != 0: tile 0 of the band is not the first tile
 * of the sequence and blends like any other.  halo_out_rect is read back as int16 (holes kept) after the last tile for the
 * next band; canvas_out as in vfsms_mosaic_host.  Either halo may be NULL. */
int vfsms_mosaic_band_host(vfsms_ctx *ctx, const uint8_t *tiles, int n_tiles, int tile_rows, int tile_cols, int channels,
                           const int32_t *tile_origin, const int32_t *roi_rect, const int32_t *pair_offset, int method,
                           int canvas_rows, int canvas_cols, int fuse_first, const int16_t *halo_in, const int32_t *halo_in_rect,
                           int16_t *halo_out, const int32_t *halo_out_rect, uint8_t *canvas_out);

/* ---------------------------------------------------------------- JPEG tile decode */

/* Replace `cv2.imdecode(np.fromfile(f, dtype=np.uint8), 0)` (Stitcher.py:68-69; the same files are decoded again at
 * :382 and :401): grayscale decode = luma component only (libjpeg JCS_GRAYSCALE), entropy decoding on the host cores
 * (files in parallel), dequantisation + accurate integer IDCT (jidctint.c islow) + range limit on the device.
 * Output is bit-identical to cv2.imdecode(..., 0).  Supported: SOF0 / SOF1 (sequential Huffman), 8-bit, one interleaved
 * scan or a single component, restart intervals; anything else -> VFSMS_E_UNSUPPORTED (callers fall back to cv2).
 * vfsms_jpeg_info and vfsms_jpeg_luma_coefficients are host-only (no context, no GPU). */
int vfsms_jpeg_info(const uint8_t *data, size_t size, int *rows, int *cols, int *components);
/* Entropy stage alone: quantised luma coefficients, int16, natural (row-major) order, blocks_h x blocks_w x 64, and the
 * luma quantisation table (natural order).  coef == NULL only reports the geometry. */
int vfsms_jpeg_luma_coefficients(const uint8_t *data, size_t size, int16_t *coef, size_t coef_capacity, int *blocks_h,
                                 int *blocks_w, uint16_t *quant /* 64 */);
/* Any component (0 = Y, 1 = Cb, 2 = Cr): same as above plus the component's sampling factors. */
int vfsms_jpeg_component_coefficients(const uint8_t *data, size_t size, int component, int16_t *coef, size_t coef_capacity,
                                      int *blocks_h, int *blocks_w, uint16_t *quant /* 64 */, int *h_samp, int *v_samp);
/* n_images files of identical geometry (rows x cols, checked).  _dev: image i lands at
 * out_dev + i * image_stride + y * row_stride + x in HBM (the device-resident tile stack the align / mosaic entry points
 * read in place).  _host: out = n_images x rows x cols u8, contiguous, host memory. */
int vfsms_jpeg_decode_gray_dev(vfsms_ctx *ctx, int n_images, const uint8_t *const *data, const size_t *sizes,
                               uint8_t *out_dev, int rows, int cols, int64_t row_stride, int64_t image_stride, void *stream);
int vfsms_jpeg_decode_gray_host(vfsms_ctx *ctx, int n_images, const uint8_t *const *data, const size_t *sizes,
                                uint8_t *out, int rows, int cols);
/* Colour decode = cv2.imdecode(data, cv2.IMREAD_COLOR) (Stitcher.py:382, 401; Main.py:6 sets isColorMode): all three
 * components through the IDCT, libjpeg's fancy chroma upsampling (jdsample.c) and YCbCr -> RGB tables (jdcolor.c), BGR
 * interleaved like cv2; single-component files give B = G = R.  Bit-identical to cv2.  out: n x rows x cols x 3;
 * _dev: pixel (y, x) of image i at out_dev + i * image_stride + y * row_stride + 3 * x. */
int vfsms_jpeg_decode_bgr_dev(vfsms_ctx *ctx, int n_images, const uint8_t *const *data, const size_t *sizes,
                              uint8_t *out_dev, int rows, int cols, int64_t row_stride, int64_t image_stride, void *stream);
int vfsms_jpeg_decode_bgr_host(vfsms_ctx *ctx, int n_images, const uint8_t *const *data, const size_t *sizes,
                               uint8_t *out, int rows, int cols);

/* Diagnostic of VFSMS_OPT_ENTROPY = 1: synchronisation passes the last device entropy decode ran (launched in rounds of 4, 8, 16, 32,
 * ...; the last round changed nothing).  Small numbers mean the subsequence decoders re-synchronised quickly. */
int vfsms_jpeg_last_entropy_passes(vfsms_ctx *ctx, int *passes_out);

/* ---------------------------------------------------------------- JPEG output encode
 * Replace `cv2.imwrite(outputAddress + ..., stitchImage)` for .jpg outputs (Stitcher.py:130-131, :196-197): baseline JPEG with
 * cv2's defaults (libjpeg quality 95 unless given, Annex-K Huffman tables, 4:2:0 for 3-channel BGR input, JFIF 1.01 header), colour
 * conversion / downsampling / integer DCT / quantisation / Huffman coding / byte stuffing all on the device; the bytes are
 * IDENTICAL to cv2.imencode(".jpg", img, [IMWRITE_JPEG_QUALITY, quality]).  channels: 1 (gray) or 3 (BGR interleaved);
 * rows, cols <= 65535.  *out_size = size of the file in bytes, also when the call returns VFSMS_E_CAPACITY because
 * out_capacity is smaller (call again with a larger buffer; out may be NULL to query). */
int vfsms_jpeg_encode_host(vfsms_ctx *ctx, const uint8_t *img, int rows, int cols, int channels, int64_t row_stride, int quality,
                           uint8_t *out, size_t out_capacity, size_t *out_size);
/* Same with the image in device memory (e.g. a mosaic canvas that never left HBM); only the compressed bytes cross PCIe. */
int vfsms_jpeg_encode_dev(vfsms_ctx *ctx, const uint8_t *img_dev, int rows, int cols, int channels, int64_t row_stride, int quality,
                          uint8_t *out, size_t out_capacity, size_t *out_size, void *stream);

/* ---------------------------------------------------------------- device-resident tile stack
 * The reference decodes every tile two or three times and ships ROIs to the plugin per call (Stitcher.py:68-69, 382, 401;
 * appendix/myGpuFeatures.cpp:70).  Here the gray tiles of a sequence (all rows x cols) are decoded / uploaded ONCE into a
 * context-owned stack in HBM; alignment reads its ROI strips in place and the gray mosaic pastes from it. */
int vfsms_tiles_reserve(vfsms_ctx *ctx, int n_tiles, int rows, int cols);
/* A second context of the same device reads (never writes) the owner's gray stack: two contexts = two streams and two
 * workspaces, so two vfsms_tiles_align* calls issued from two host threads overlap on the GPU (the row-strip and the
 * column-strip candidates of one search round, sharding._run_requests).  The owner must outlive the borrower's use and must
 * not re-reserve meanwhile; uploads / decodes go through the owner only.  Synchronises the owner's stream. */
int vfsms_tiles_attach(vfsms_ctx *ctx, vfsms_ctx *owner);
int vfsms_tiles_decode_jpeg(vfsms_ctx *ctx, int first, int n, const uint8_t *const *data, const size_t *sizes);   /* as vfsms_jpeg_decode_gray_dev */
int vfsms_tiles_upload(vfsms_ctx *ctx, int first, int n, const uint8_t *tiles /* n x rows x cols, host */);
int vfsms_tiles_download(vfsms_ctx *ctx, int first, int n, uint8_t *out /* n x rows x cols, host */);
const uint8_t *vfsms_tiles_ptr(vfsms_ctx *ctx);                      /* device address of tile 0 */
/* One search candidate (ROI length roi_len = floor(edge * i * roiRatio), direction 1..4) of calculateOffsetForFeatureSearchIncre
 * (Stitcher.py:319-351) for the n_pairs consecutive pairs (first + p, first + p + 1): ROIs as getROIRegionForIncreMethod
 * cuts them (ImageUtility.py:66-101), read in place; results as vfsms_align_batch_host (ROI coordinates, host memory). */
int vfsms_tiles_align(vfsms_ctx *ctx, int first, int n_pairs, int direction, int roi_len, const vfsms_surf_params *params,
                      float ratio, int offset_evaluate, vfsms_pair_result *results);
/* The same for every pair_step-th pair: pairs (first + p * pair_step, first + p * pair_step + 1), p < n_pairs -- one fused call for
 * the probe pairs with which a sharded run guesses the shooting direction before it walks its pairs (sharding.py). */
int vfsms_tiles_align_strided(vfsms_ctx *ctx, int first, int n_pairs, int pair_step, int direction, int roi_len,
                              const vfsms_surf_params *params, float ratio, int offset_evaluate, vfsms_pair_result *results);
/* The same for an arbitrary list: pair p = tiles (first_tile[p], first_tile[p] + 1) in direction[p].  All directions of one call
 * must cut strips of one shape (1 and 3: roi_len x cols, 2 and 4: rows x roi_len).  One fused call for a whole round of the batched
 * search of a sharded run, whatever pairs and directions it asks for (sharding.evaluate_shard_batched). */
int vfsms_tiles_align_list(vfsms_ctx *ctx, int n_pairs, const int32_t *first_tile, const int32_t *direction, int roi_len,
                           const vfsms_surf_params *params, float ratio, int offset_evaluate, vfsms_pair_result *results);
/* vfsms_mosaic_host with tiles first .. first + n_tiles - 1 of the stack (gray). */
int vfsms_tiles_mosaic(vfsms_ctx *ctx, int first, int n_tiles, const int32_t *tile_origin, const int32_t *roi_rect,
                       const int32_t *pair_offset, int method, int canvas_rows, int canvas_cols, uint8_t *canvas_out);

/* Colour twin of the stack.  Main.py:6 sets isColorMode = True: the reference then decodes every file in gray for the alignment
 * (Stitcher.py:68-69) and again in colour for the mosaic (Stitcher.py:382, :401).  vfsms_tiles_decode_jpeg_bgr entropy-decodes
 * each file ONCE: BGR (= cv2.imdecode(data, IMREAD_COLOR)) goes to the colour twin and the luma plane (= cv2.imdecode(data, 0),
 * libjpeg's JCS_GRAYSCALE output) to the gray stack, slots first .. first + n - 1.  vfsms_tiles_upload_bgr fills colour slots from
 * host memory (other formats; the gray slot is filled by vfsms_tiles_upload).  vfsms_tiles_mosaic_bgr = vfsms_mosaic_host with
 * channels = 3 on tiles that never left HBM; every slot of the range must hold a colour tile. */
int vfsms_tiles_decode_jpeg_bgr(vfsms_ctx *ctx, int first, int n, const uint8_t *const *data, const size_t *sizes);
int vfsms_tiles_upload_bgr(vfsms_ctx *ctx, int first, int n, const uint8_t *tiles_bgr /* n x rows x cols x 3, host */);
int vfsms_tiles_mosaic_bgr(vfsms_ctx *ctx, int first, int n_tiles, const int32_t *tile_origin, const int32_t *roi_rect,
                           const int32_t *pair_offset, int method, int canvas_rows, int canvas_cols, uint8_t *canvas_out /* rows x cols x 3 */);

#ifdef __cplusplus
}
#endif
#endif /* VFSMS_H */
