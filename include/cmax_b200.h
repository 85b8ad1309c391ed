/* cmax_b200.h - C ABI of the B200-native contrast-maximisation (CMax) loss path.
 *
 * Drop-in boundary for tub-rip/MotionPriorCMax's loss plugin (upstream file:line cited per
 * entry point).  Plain pointers and sizes only: no torch types, no C++ in the signatures.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless it says "host"; all tensors are dense,
 *     row-major, float32 unless stated; the caller owns every buffer;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     enqueued on it, nothing synchronises, nothing allocates: scratch comes from the caller's
 *     `workspace` (size from cmax_workspace_bytes, 256-byte aligned base);
 *   - return value: CMAX_OK (0) or a negative CMAX_ERR_* code; never throws;
 *   - coordinates are (y, x); events are rows of 6 floats (y, x, t, p, bin, valid) with the
 *     positives first when polarity-aware (upstream src/loader/dsec/loader.py:156-161,360-415).
 */
#ifndef CMAX_B200_H_
#define CMAX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMAX_ABI_VERSION 1

enum {
    CMAX_OK = 0,
    CMAX_ERR_BAD_CONFIG = -1,   /* a CmaxConfig field is out of range / forbidden combination */
    CMAX_ERR_BAD_SHAPE = -2,    /* B, M, n, num_pos_events inconsistent                          */
    CMAX_ERR_WORKSPACE = -3,    /* workspace NULL, misaligned or smaller than required           */
    CMAX_ERR_CUDA = -4,         /* a CUDA launch failed (cudaGetLastError is left set)           */
    CMAX_ERR_UNSUPPORTED = -5   /* valid but not implemented size (e.g. num_knn too large)       */
};

/* norms / schemes */
enum { CMAX_NORM_L1 = 0, CMAX_NORM_L2 = 1 };
enum { CMAX_INTERP_MEAN = 0, CMAX_INTERP_IWD = 1 };
enum { CMAX_SMOOTH_ON_FLOW_TO_TREF = 0, CMAX_SMOOTH_ON_FLOW_TO_NEXT = 1 };
enum { CMAX_BASIS_POLYNOMIAL = 0, CMAX_BASIS_DCT = 1, CMAX_BASIS_BEZIER = 2 };
enum { CMAX_FOCUS_GRADIENT_MAGNITUDE = 0, CMAX_FOCUS_VARIANCE = 1 };

/* Mirrors the keyword arguments of upstream FocusLoss.__init__ (src/losses/focus.py:28-51). */
typedef struct CmaxConfig {
    int32_t height, width;            /* image_shape                                        */
    int32_t num_tref;                 /* R >= 1                                             */
    int32_t num_bins;                 /* nb >= 1                                            */
    int32_t num_knn;                  /* K >= 1, K <= n                                     */
    int32_t lut_superpixel_size;      /* s >= 1                                             */
    int32_t focus_loss_norm;          /* CMAX_NORM_L1 | CMAX_NORM_L2                        */
    int32_t dist_norm;                /* CMAX_NORM_L1 | CMAX_NORM_L2                        */
    int32_t scale_iwe_by_dt;          /* bool; requires num_tref == 1 (focus.py:49)         */
    int32_t mask_image_border;        /* bool                                               */
    int32_t polarity_aware_batching;  /* bool; requires num_tref == 1 (focus.py:50)         */
    int32_t interpolation_scheme;     /* CMAX_INTERP_MEAN | CMAX_INTERP_IWD                 */
    int32_t smooth_type;              /* on_flow_to_next requires num_tref == 1 (focus.py:51) */
    float smooth_weight;
    int32_t deterministic;            /* 1: IWE and LUT-gradient accumulate in int64 fixed point
                                         (run-to-run bit-identical); 0: float32 atomics          */
    int32_t focus_functional;         /* CMAX_FOCUS_GRADIENT_MAGNITUDE (what upstream calc hard-codes,
                                         focus.py:90) | CMAX_FOCUS_VARIANCE (loss.py:14-16)       */
    int32_t backward_follows;         /* training hint: 1 = cmax_backward will be called on this
                                         workspace.  cmax_forward then emits dL/dIWE from the same pass
                                         that blurs / scores the IWE (one kernel instead of two; the
                                         matching cmax_backward, given the SAME config, skips that
                                         stage).  Results are unchanged.  0 = forward only / unknown. */
    int32_t reserved[1];
} CmaxConfig;

int cmax_abi_version(void);
const char *cmax_error_string(int code);

/* Scratch bytes cmax_forward / cmax_backward need for a batch of B windows of M event rows and
 * n trajectories.  Returns 0 on an invalid configuration.  The same workspace must be handed,
 * untouched, from cmax_forward to the matching cmax_backward (it carries the saved context). */
size_t cmax_workspace_bytes(const CmaxConfig *cfg, int64_t B, int64_t M, int64_t n);

/* Forward of upstream FocusLoss.calc (src/losses/focus.py:66-113): interpolate_flow (:115-180),
 * warp_events (:182-195), make_iwes (:197-230) -> create_iwe (src/utils/event_image_converter.py
 * :45-74,333-391 + gaussian blur :170-175), calculate_focus_loss (src/utils/loss.py:4-27) and
 * calculate_smooth_loss (focus.py:232-246, loss.py:29-56).
 *
 *   trajectories [B, R + nb, n, 2]   times [R + nb]   events [B, M, 6]
 *   num_pos_events: split index of the polarity-aware layout (ignored otherwise; must be
 *                   0..M when polarity_aware_batching)
 *   iwes_out   [B * R, P, H, W]  blurred IWEs, P = 2 if polarity aware else 1
 *   losses_out [3] = {loss, focus_loss, smoothness_loss}
 *   flow_lut_out (optional, may be NULL) [B, nb, H/s, W/s, R, 2]
 */
int cmax_forward(const CmaxConfig *cfg, const float *trajectories, const float *times,
                 const float *events, int64_t B, int64_t M, int64_t n, int64_t num_pos_events,
                 float *iwes_out, float *losses_out, float *flow_lut_out,
                 void *workspace, size_t workspace_bytes, void *stream);

/* Backward of the same call: d loss / d trajectories, scaled by *grad_loss (device scalar).
 *   dtraj_out [B, R + nb, n, 2] is fully overwritten.
 * Inputs must be the ones given to the cmax_forward that filled `workspace`. */
int cmax_backward(const CmaxConfig *cfg, const float *trajectories, const float *times,
                  const float *events, int64_t B, int64_t M, int64_t n, int64_t num_pos_events,
                  const float *grad_loss, float *dtraj_out,
                  void *workspace, size_t workspace_bytes, void *stream);

/* ---- packed, tile-binned event layout (the loader-side layout of SURVEY.md 8f rank 2) ----------
 * Replaces the `batch['events']` tensor the reference collate builds (src/loader/dsec/loader.py
 * :141-182,360-415) by 16-byte records without padding rows, grouped so that the event kernels can
 * accumulate in shared memory:
 *   records   [B, M, 4] float32: (y, x, t, meta), meta = the bit pattern of
 *             bin << 24 | iy << 12 | ix, the event's LUT cell (upstream focus.py:185-187:
 *             it = int(bin), iy = int(y // s), ix = int(x // s)); only rows with valid != 0 whose
 *             cell lies inside the table are kept (valid is treated as 1);
 *   seg_start [B, G * NT + 1] int32: per sample, prefix offsets of the segments ordered by
 *             (polarity group g < G, source tile T < NT); G = 2 when polarity aware, else 1;
 *             tile T = (iy / ct) * tiles_x + ix / ct; the last entry is the number of records.
 * The order of the records inside a segment is free.
 * cmax_pack_layout: host query, out = {ct (LUT cells per tile edge), tiles_y, tiles_x, G}.
 * cmax_pack_events: builds the layout on the device from an upstream-layout events tensor;
 *   scratch [B, G * NT] int32; skipped_out (optional) int64[2] DEVICE counters:
 *   [0] valid rows dropped because their LUT cell is outside the table (cmax_forward counts the
 *   same rows in its status word), [1] rows whose `valid` is neither 0 nor 1.
 * Limits: num_bins <= 256, H/s and W/s <= 4096 (else CMAX_ERR_UNSUPPORTED). */
int cmax_pack_layout(const CmaxConfig *cfg, int32_t out_host[4]);
int cmax_pack_events(const CmaxConfig *cfg, const float *events, int64_t B, int64_t M,
                     int64_t num_pos_events, float *records_out, int32_t *seg_start_out,
                     int32_t *scratch, int64_t *skipped_out, void *stream);

/* HOST builder of the same layout for DataLoader workers (C++ / OpenMP over the windows, no CUDA
 * call): all pointers are HOST pointers.  A stable counting sort - the row order inside every
 * segment is kept, so the result equals io.pack_events_host byte for byte.
 *   records_host [B, records_stride, 4] (may be NULL: a first call that only fills seg_start tells
 *   the caller how many records each window needs, seg_start[b, G*NT]); skipped_host int64[2]. */
int cmax_pack_events_host(const CmaxConfig *cfg, const float *events_host, int64_t B, int64_t M,
                          int64_t num_pos_events, float *records_host, int64_t records_stride,
                          int32_t *seg_start_host, int64_t *skipped_host);

/* Compact WIRE layout for the host -> device copy (12 bytes per valid event instead of 24 / 16):
 *   coords     [T, 3] float32 (y, x, t): the windows of the batch back to back, no padding rows;
 *              window b owns rows sample_off[b] .. sample_off[b + 1] (int64 [B + 1]);
 *   fine_start [B, G * NT * nb + 1] int32: per window, prefix offsets (relative to its first row)
 *              of the runs ordered by (polarity group, source tile, time bin): the bin column of
 *              the reference layout (focus.py:185) is implied by the run, the LUT cell by (y, x).
 * cmax_pack_events_host_compact (HOST pointers, C++ / OpenMP, for the loader workers): call once
 *   with coords_host = NULL to get fine_start and sample_off (sizes), then with a buffer of at
 *   least sample_off[B] rows.  cmax_expand_compact (DEVICE pointers): one kernel rebuilds the
 *   16-byte records [B, records_stride, 4] and seg_start [B, G * NT + 1] of the packed layout,
 *   records_stride >= the largest window; LUT cells recomputed with the reference's arithmetic. */
int cmax_pack_events_host_compact(const CmaxConfig *cfg, const float *events_host, int64_t B, int64_t M,
                                  int64_t num_pos_events, float *coords_host, int64_t coords_capacity,
                                  int32_t *fine_start_host, int64_t *sample_off_host,
                                  int64_t *skipped_host);
int cmax_expand_compact(const CmaxConfig *cfg, const float *coords, const int32_t *fine_start,
                        const int64_t *sample_off, int64_t B, int64_t records_stride,
                        float *records_out, int32_t *seg_start_out, int32_t *scratch, void *stream);
/* int32 elements of DEVICE scratch both expand calls need (the run of the first record of every
 * block of 256 records, filled by a tiny kernel before the expansion). */
int64_t cmax_expand_scratch_ints(const CmaxConfig *cfg, int64_t B, int64_t records_stride);

/* Bit-packed wire layout: the runs of the compact layout, every run stored as fixed-width records of
 * bit-pattern deltas (inside one (group, tile, bin) run y, x and t vary little, and the IEEE bit
 * pattern of a non-negative float is monotone): field = bits(value) - min over the run, in
 * w = bit_length(max - min) bits, w = 0..32 per field.  LOSSLESS for any float32 input; about 8 B per
 * event for a DSEC window instead of 12.
 *   words    [total] uint32: the windows' bit streams back to back, window b at word_off[b]
 *            (int64 [B + 1]; two zero words of slack end every window's stream);
 *   run_hdr  [B, F, 4] uint32 (F = G * NT * nb): min patterns of y, x, t and wy | wx << 8 | wt << 16;
 *   run_word [B, F + 1] int32: first word of every run inside its window's stream (runs are word
 *            aligned); fine_start [B, F + 1] as in the compact layout.
 * cmax_pack_events_host_bitpacked (HOST pointers): words_host = NULL fills only the tables (word_off =
 * the sizes); with words_host it fills tables and streams in the same call and returns
 * CMAX_ERR_BAD_SHAPE (tables filled in) when words_capacity is too small - B * (3 * M + F + 2) words
 * always suffice,
 * cmax_expand_bitpacked (DEVICE pointers): decodes into the packed layout, bit for bit the records
 * cmax_expand_compact produces. */
int cmax_pack_events_host_bitpacked(const CmaxConfig *cfg, const float *events_host, int64_t B, int64_t M,
                                    int64_t num_pos_events, uint32_t *words_host, int64_t words_capacity,
                                    int32_t *fine_start_host, uint32_t *run_hdr_host, int32_t *run_word_host,
                                    int64_t *word_off_host, int64_t *skipped_host);
int cmax_expand_bitpacked(const CmaxConfig *cfg, const uint32_t *words, const int32_t *fine_start,
                          const uint32_t *run_hdr, const int32_t *run_word, const int64_t *word_off, int64_t B,
                          int64_t records_stride, float *records_out, int32_t *seg_start_out, int32_t *scratch,
                          void *stream);

/* cmax_forward / cmax_backward on the packed layout: same outputs, same workspace
 * (cmax_workspace_bytes(cfg, B, M, n) with the M of `records`).  The event stage accumulates
 * the IWE votes of a (tile, group) segment in a shared-memory window and flushes once; votes
 * leaving the window fall back to global atomics, so the result is exact for any flow. */
int cmax_forward_packed(const CmaxConfig *cfg, const float *trajectories, const float *times,
                        const float *records, const int32_t *seg_start, int64_t B, int64_t M,
                        int64_t n, float *iwes_out, float *losses_out, float *flow_lut_out,
                        void *workspace, size_t workspace_bytes, void *stream);
int cmax_backward_packed(const CmaxConfig *cfg, const float *trajectories, const float *times,
                         const float *records, const int32_t *seg_start, int64_t B, int64_t M,
                         int64_t n, const float *grad_loss, float *dtraj_out, void *workspace,
                         size_t workspace_bytes, void *stream);

/* ---- phased calls: one window (or batch) whose EVENT ROWS are sharded over ranks -------------
 * (second mode).  Every rank holds the same trajectories and its own slice of the
 * event rows (any M per rank).  cmax_forward_accumulate builds the LUT and splats the local
 * events into the raw-IWE section of the workspace; the caller sums that section over the ranks
 * (ncclAllReduce / torch.distributed.all_reduce, in place); cmax_forward_finish runs the image
 * stage on the summed IWE.  Backward: cmax_backward_accumulate leaves the local part of
 * d loss / d LUT in the dLUT section (pass include_smooth = 1 on exactly one rank so that the
 * smoothness gradient enters the sum once; ignored when deterministic), the caller all-reduces
 * it, cmax_backward_finish gathers it into the trajectories.  Every rank ends with the same loss,
 * IWEs and gradients as a single cmax_forward / cmax_backward over all the rows; in deterministic
 * mode (int64 sections) bit for bit.
 * cmax_workspace_section: byte offset / size of a reducible section inside the workspace and
 * whether it holds int64 (deterministic) or float32 values. */
enum { CMAX_SECTION_RAW_IWE = 0, CMAX_SECTION_DLUT = 1 };
int cmax_workspace_section(const CmaxConfig *cfg, int64_t B, int64_t M, int64_t n, int32_t which,
                           size_t *offset_out, size_t *bytes_out, int32_t *is_int64_out);
int cmax_forward_accumulate(const CmaxConfig *cfg, const float *trajectories, const float *times,
                            const float *events, int64_t B, int64_t M, int64_t n,
                            int64_t num_pos_events, float *flow_lut_out, void *workspace,
                            size_t workspace_bytes, void *stream);
int cmax_forward_finish(const CmaxConfig *cfg, int64_t B, int64_t M, int64_t n, float *iwes_out,
                        float *losses_out, void *workspace, size_t workspace_bytes, void *stream);
int cmax_backward_accumulate(const CmaxConfig *cfg, const float *trajectories, const float *times,
                             const float *events, int64_t B, int64_t M, int64_t n,
                             int64_t num_pos_events, const float *grad_loss, int32_t include_smooth,
                             void *workspace, size_t workspace_bytes, void *stream);
int cmax_backward_finish(const CmaxConfig *cfg, const float *trajectories, int64_t B, int64_t M,
                         int64_t n, const float *grad_loss, float *dtraj_out, void *workspace,
                         size_t workspace_bytes, void *stream);

/* Stand-alone imager, upstream EventImageConverter.create_iwe(events, method='bilinear_vote',
 * sigma, weight) (src/utils/event_image_converter.py:45-74,134-176,333-391).
 *   events [nb, M, row_stride] (first two columns y, x; row_stride >= 2 floats)
 *   weight NULL (-> 1.0) or [nb, M];  sigma 0 (no blur) or > 0 (3x3 gaussian, reflect pad)
 *   out [nb, H, W];  scratch: NULL when sigma == 0, else nb*H*W floats.
 *   deterministic: accumulate in int64 fixed point (scratch_i64 [nb*H*W] required). */
int cmax_create_iwe(const float *events, const float *weight, int64_t nb, int64_t M,
                    int64_t row_stride, int32_t H, int32_t W, float sigma, float *out,
                    float *scratch, int64_t *scratch_i64, int32_t deterministic, void *stream);

/* Event-count image, upstream count_event_tensor (event_image_converter.py:226-272): four unit
 * votes per event at the bilinear corners, exact integers.  out [nb, H, W] int64. */
int cmax_count_image(const float *events, int64_t nb, int64_t M, int64_t row_stride,
                     int32_t H, int32_t W, int64_t *out, void *stream);

/* The same two with the imager's outer_padding (event_image_converter.py:23-28, 339-343): H, W are
 * the PADDED image sizes (image + 2 * pad), the integer corner of every event is shifted by
 * (pad_y, pad_x) after the floor; fractional parts and weights are unchanged.  pad = 0 is the
 * plain call. */
int cmax_create_iwe_padded(const float *events, const float *weight, int64_t nb, int64_t M,
                           int64_t row_stride, int32_t H, int32_t W, int32_t pad_y, int32_t pad_x,
                           float sigma, float *out, float *scratch, int64_t *scratch_i64,
                           int32_t deterministic, void *stream);
int cmax_count_image_padded(const float *events, int64_t nb, int64_t M, int64_t row_stride,
                            int32_t H, int32_t W, int32_t pad_y, int32_t pad_x, int64_t *out,
                            void *stream);

/* Test / inspection entry: the K nearest trajectories of every LUT cell exactly as cmax_forward
 * selects them (focus.py:128-137), ascending distance, lowest index first on ties.
 *   points [S, n, 2] (S independent slabs = B * nb);  ind_out [S, q, K] int32;
 *   dist_out (optional) [S, q, K].  workspace: cmax_knn_workspace_bytes. */
size_t cmax_knn_workspace_bytes(int32_t H, int32_t W, int32_t s, int64_t S, int64_t n, int32_t K);
int cmax_knn_indices(const float *points, int64_t S, int64_t n, int32_t H, int32_t W, int32_t s,
                     int32_t K, int32_t dist_norm, int32_t *ind_out, float *dist_out,
                     void *workspace, size_t workspace_bytes, void *stream);

/* Fused front end, upstream TrajectoryNet.calculate_trajectories_at_t
 * (src/modules/trajectory_net.py:101-119) = coeffs_grid_to_list (src/utils/trajectories.py:15-52)
 * + compute_basis (src/utils/basis.py:4-46), and the Bezier basis of
 * src/models/raft_spline/curves/bezier.py:68-113.
 *   coeff_grid [B, S, 2K, H, W]; channel c < K -> first axis, c >= K -> second axis; axis order
 *   (y, x) unless xy_order != 0 (RAFT-spline stores x first);  basis_table host-independent:
 *   phi [n_t, K] DEVICE table of (phi_k(t) - phi_k(anchor)) built by the caller;
 *   trajectories_out [B, n_t, n, 2] with n = ceil((H - p/2)/p) * ceil((W - p/2)/p) tile centres. */
int cmax_trajectories_forward(const float *coeff_grid, const float *phi, int64_t B, int64_t S,
                              int32_t K, int32_t H, int32_t W, int32_t patch, int32_t n_t,
                              int32_t xy_order, int32_t add_offsets, float *trajectories_out,
                              void *stream);
/* Adjoint: d coeff_grid [B, S, 2K, H, W] (dense, zero away from the tile centres). */
int cmax_trajectories_backward(const float *dtraj, const float *phi, int64_t B, int64_t S,
                               int32_t K, int32_t H, int32_t W, int32_t patch, int32_t n_t,
                               int32_t xy_order, float *dcoeff_grid_out, void *stream);

/* Event voxel grid (the network input), upstream VoxelGrid.convert (src/loader/dsec/utils.py:29-77):
 * trilinear vote of 2p-1 into grid_out [C, H, W] with t_norm = (C-1)(t - t[0])/(t[n-1] - t[0]),
 * then norm_type 0: none, 1: 'mean_std' over the non-zero entries (unbiased std), 2: 'max'.
 *   x, y, t, p [n] float32 (t sorted);  stats_scratch: 4 doubles (required when norm_type != 0). */
int cmax_voxel_grid(const float *x, const float *y, const float *t, const float *p, int64_t n,
                    int32_t C, int32_t H, int32_t W, int32_t norm_type, float *grid_out,
                    double *stats_scratch, void *stream);
/* Second half of VoxelGrid.convert on an existing grid (utils.py:56-75), in place: optional
 * quantile clipping |v| > thr -> sign(v) * thr with thr = *clip_threshold (DEVICE pointer, NULL =
 * no clipping; the order statistic itself is the caller's), then the normalisation as above. */
int cmax_voxel_normalize(float *grid, int32_t C, int32_t H, int32_t W, int32_t norm_type,
                         const float *clip_threshold, double *stats_scratch, void *stream);

/* Dense flow read-out, upstream dense_flow_from_traj (src/utils/flow.py:8-16): list_to_grid
 * (src/utils/trajectories.py:54-75) on the patch lattice, then torchvision BICUBIC antialias
 * resize to the image.  traj_flow [B, n, C], pixel_positions [n, 2] int64 (y, x);
 * patch_flow_out [B, C, H/patch, W/patch], dense_out [B, C, H, W].  Forward only. */
int cmax_dense_flow(const float *traj_flow, const int64_t *pixel_positions, int64_t B, int64_t n,
                    int32_t C, int32_t patch, int32_t H, int32_t W, float *patch_flow_out,
                    float *dense_out, void *stream);

/* Micro-benchmark used by bench.py to measure the atomic side of the roofline on the box:
 * n_ops float32 `red.global.add` to pseudo-random addresses inside `region_floats` floats.
 * mode 0: global red.f32, 1: shared-memory atomics + flush, 2: global red on int64,
 * 3: red.global.add.v2.f32 (n_ops / 2 requests: the two x-adjacent corners of a vote per request),
 * 4: red.global.add.v4.f32 (n_ops / 2 requests: the corner pair in the middle of an aligned quad). */
int cmax_atomic_microbench(float *region, int64_t region_floats, int64_t n_ops, int32_t mode,
                           void *stream);

/* Measurement hooks (bench.py / profiles): per-stage device time from cudaEvent pairs recorded on
 * the launching stream around every kernel of cmax_forward / cmax_backward / the front end.
 * enable(1) starts recording (up to 8192 stage launches), read() synchronises on the recorded
 * events, returns the summed milliseconds and launch counts per stage and clears the buffer.
 * Not thread safe (one caller thread per process, as in the reference). */
int cmax_stage_count(void);
const char *cmax_stage_name(int stage);
int cmax_stage_timing_enable(int on);
int cmax_stage_timing_read(double *ms_sum /* [cmax_stage_count()] */, int64_t *count);
/* Number of kernels this library has launched since it was loaded. */
int64_t cmax_launch_count(void);
/* Inspection: how many LUT queries of the most recent cmax_forward / cmax_knn_indices call left
 * the staged fast path for the heap search (host call, synchronises `stream`; -1 if unknown). */
int64_t cmax_last_worklist_count(void *stream);
/* Same call, split by reason: out[0] = cells that left the staged kernel, out[1..7] = window not
 * staged (too many points), no usable bracket, fewer than K / more than 255 points in the window,
 * K-th key beyond the window's guaranteed bound, boundary list full, previous-bin bracket missed;
 * out[8] = cells the warp-cooperative work-list kernel passed on to the heap search; rest reserved. */
int cmax_last_worklist_reasons(int64_t out_host[16], void *stream);

/* Reads the status words the kernels keep in the workspace (host call, synchronises `stream`):
 * out[0] = events skipped because their LUT cell index was out of range, out[1..3] reserved. */
int cmax_read_status(const void *workspace, int64_t out_host[4], void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CMAX_B200_H_ */
