// cmax_common.cuh - shared declarations of the B200 (sm_100a) CMax loss kernels.
//
// Data layout in HBM (all float32 unless noted, row-major):
//   trajectories [B, R+nb, n, 2]      (y, x) absolute pixel positions, caller owned
//   events       [B, M, 6]            (y, x, t, p, bin, valid), caller owned
//   workspace    one caller-owned slab carved by `Layout` below: LUT, per-cell KNN
//                thresholds, raw IWE, dL/dIWE, dLUT ...
// "slab" = one (sample, time-bin) pair: S = B * nb independent KNN problems.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/cmax_b200.h"

namespace cmax {

constexpr int kMaxCells = 24576;      // cell-list cells per slab (smem counters, 96 KB)
constexpr int kMaxTref = 16;          // reference times handled by the gather kernels
constexpr int kKnnBlock = 128;        // threads per KNN CTA (one LUT query per thread)
#ifndef CMAX_KNN_TILE_W
#define CMAX_KNN_TILE_W 16
#endif
constexpr int kKnnTileW = CMAX_KNN_TILE_W;         // KNN CTA = 16 x 8 queries
constexpr int kKnnTileH = kKnnBlock / kKnnTileW;
constexpr int kMaxKnn = 192;          // heap lives in shared memory: 8 B * K * 128 threads
constexpr int kImgTile = 32;          // image-stage CTA tile (32 x 32 pixels, 256 threads)
constexpr double kFixScale = 4294967296.0;   // 2^32: int64 fixed-point scale (deterministic mode)
constexpr float kVoteEps = 1e-6f;     // event_image_converter.py:357
constexpr float kIwdEps = 1e-9f;      // focus.py:7
constexpr float kCharbEps2 = 1e-6f;   // (1e-3)^2, loss.py:46,55

struct Geom {                 // derived sizes, passed by value to kernels
    int H, W, R, nb, K, s, P;
    int Hq, Wq, q;            // LUT lattice
    float off;                // s/2 - 0.5 (focus.py:117)
    int Hc, Wc, NC;           // cell list
    float cs, inv_cs;         // cell edge (multiple of s)
    int r0;                   // initial search radius in cells (heap path)
    int r_fast;               // window radius of the staged fast path
    int l1dist, l2focus, scale_dt, mask_border, pab, iwd, smooth_next, det, variance;
    int fuse_image;           // training hint: the forward also produces dL/dIWE (image stage fused)
    float smooth_w;
    int64_t B, M, n, S;       // S = B * nb
    int64_t npos;
    int ct, nty, ntx, nt;     // packed event layout: source tiles of ct x ct LUT cells (tile_stage.cu)
};

struct Header {               // first 1 KiB of the workspace
    long long status[4];      // [0] events skipped: LUT cell out of range
    double focus_sum;         // sum of |dx|+|dy| (or squares) over all IWE pixels
    double smooth_sum;        // sum of charbonnier terms (x and y)
    float val;                // focus_sum / N
    float focus, smooth, loss;
};

struct Layout {
    size_t header, focus_partials, plane_stats, smooth_partials, cell_start, sorted, recs, sorted_j, tau, jcut, wsum,
        tau_max, tile_max, worklist, worklist2, work_count, bpart, lut, f2n, raw, raw_i64, dimg, dlut, dlut_i64, df2n, total;
    int n_img_blocks, n_sm_blocks;
};

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

int make_geom(const CmaxConfig *cfg, int64_t B, int64_t M, int64_t n, int64_t npos, Geom *g);
void knn_geom(int H, int W, int s, int64_t n, int K, Geom *g);
Layout make_layout(const Geom &g);
Layout make_knn_layout(const Geom &g);     // only the fields the KNN kernels touch

// ---- launchers (each returns cudaGetLastError() != cudaSuccess ? CMAX_ERR_CUDA : CMAX_OK) ----
int launch_lut_forward(const Geom &g, const Layout &L, const float *traj, char *ws,
                       float *flow_lut_out, int32_t *ind_out, float *dist_out, cudaStream_t st);
int launch_lut_backward(const Geom &g, const Layout &L, const float *traj, char *ws,
                        float *dtraj, cudaStream_t st);
int launch_event_forward(const Geom &g, const Layout &L, const float *events, const float *times,
                         char *ws, cudaStream_t st, int phase = 0);
int launch_event_backward(const Geom &g, const Layout &L, const float *events, const float *times,
                          const float *grad_loss, char *ws, cudaStream_t st, int phase = 0);
int launch_pack_events(const Geom &g, const float *events, float4 *records, int *seg_start,
                       int *scratch, long long *skipped, cudaStream_t st);
int launch_expand_compact(const Geom &g, const float *coords, const int *fine_start,
                          const long long *sample_off, int64_t Mp, float4 *records, int *seg_start,
                          int *block_run, cudaStream_t st);
int launch_expand_bitpacked(const Geom &g, const unsigned *words, const int *fine_start, const unsigned *run_hdr,
                            const int *run_word, const long long *word_off, int64_t Mp, float4 *records,
                            int *seg_start, int *block_run, cudaStream_t st);
int launch_event_forward_packed(const Geom &g, const Layout &L, const float4 *records,
                                const int *seg_start, const float *times, char *ws, cudaStream_t st);
int launch_event_backward_packed(const Geom &g, const Layout &L, const float4 *records,
                                 const int *seg_start, const float *times, const float *grad_loss,
                                 char *ws, cudaStream_t st);
int launch_fix_to_float(const long long *in, float *out, int64_t count, cudaStream_t st);
int launch_dlut_finalize(const Geom &g, const Layout &L, const float *grad_loss, char *ws,
                         cudaStream_t st);
int launch_image_forward(const Geom &g, const Layout &L, char *ws, float *iwes_out, cudaStream_t st);
int launch_image_backward(const Geom &g, const Layout &L, char *ws, cudaStream_t st);
int launch_image_forward_backward(const Geom &g, const Layout &L, char *ws, float *iwes_out,
                                  cudaStream_t st);
int launch_smooth_forward(const Geom &g, const Layout &L, char *ws, cudaStream_t st);
int launch_smooth_backward(const Geom &g, const Layout &L, const float *grad_loss, char *ws,
                           cudaStream_t st);
int launch_finalize_losses(const Geom &g, const Layout &L, char *ws, float *losses_out,
                           cudaStream_t st);

// Launches go to the *current* device: make that the device the caller's buffers live on (tensors of
// a non-current GPU in a single process), restore on exit.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(const void *p)
    {
        cudaPointerAttributes a;
        int cur = 0;
        if (p && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice &&
            cudaGetDevice(&cur) == cudaSuccess && a.device != cur) {
            if (cudaSetDevice(a.device) == cudaSuccess) prev = cur;
        }
        cudaGetLastError();      // a host pointer / NULL is not an error here
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

inline int check_launch() { return cudaGetLastError() == cudaSuccess ? CMAX_OK : CMAX_ERR_CUDA; }

// ---- optional per-stage timing (cudaEvent pairs on the launching stream) and launch counter ----
enum Stage {
    ST_BIN_POINTS = 0, ST_KNN_SELECT, ST_EVENT_FWD, ST_IMAGE_FWD, ST_SMOOTH_FWD, ST_FINALIZE,
    ST_IMAGE_BWD, ST_SMOOTH_BWD, ST_EVENT_BWD, ST_LUT_BWD, ST_TRAJ_FWD, ST_TRAJ_BWD, ST_PACK, ST_COUNT
};
void stage_begin(int stage, cudaStream_t st);
void stage_end(int stage, cudaStream_t st);
void count_launch(int n = 1);
struct StageScope {
    int id;
    cudaStream_t st;
    StageScope(int i, cudaStream_t s) : id(i), st(s) { stage_begin(id, st); }
    ~StageScope() { stage_end(id, st); }
};

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
// 8-byte read-only load of the event stream.  A row is 24 B, so the three loads of a warp touch
// the same 128 B lines: letting them allocate in L1 (measured: -3 % event_forward, -2 %
// event_backward versus L1::no_allocate, which re-fetches the shared sectors from L2 three times).
__device__ __forceinline__ float2 ld_stream_f2(const float *p)
{
    float2 r;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

// Two adjacent floats in one L2 atomic request (sm_90+): halves the red traffic of (gy, gx)
// pairs and of x-adjacent bilinear corners.  `p` must be 8-byte aligned.
__device__ __forceinline__ void red_add_f32x2(float *p, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// Four adjacent floats in one L2 atomic request; `p` must be 16-byte aligned.
__device__ __forceinline__ void red_add_f32x4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// v * 2^32 as int64 (deterministic fixed point): the product is exact in float32 (power of two),
// so this equals llrint((double)v * 2^32) bit for bit - without FP64 arithmetic.
__device__ __forceinline__ long long to_fix(float v)
{
    return __float2ll_rn(__fmul_rn(v, 4294967296.0f));
}

// ---- bulk async copies (TMA engine, no tensor map): global -> shared, completion on an mbarrier ----
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(parity)
                 : "memory");
}

// sqrt.approx (MUFU, ~1 ulp): ONLY for search bounds that carry their own safety margin - never
// for a value the reference computes.  The IEEE sqrtf costs ~15 instructions under -prec-sqrt.
__device__ __forceinline__ float approx_sqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Python-style float floor division, the arithmetic of torch's `//` on float tensors
// (c10 div_floor_floating) used by focus.py:186-187.
__device__ __forceinline__ float floordiv_f32_generic(float a, float b);

// Fast path of the same function for a divisor that is a power of two (every shipped config:
// lut_superpixel_size = 4): fmod is exact, a - mod is an exact multiple of b and the quotient an
// integer, so the whole recipe below reduces to floor(a / b) - and a * (1 / b) is exact as long as
// it does not underflow (|a| >= 1e-30; smaller magnitudes take the generic recipe).  Saves two
// ~40-instruction fmodf sequences per event.
__device__ __forceinline__ float floordiv_f32(float a, float b)
{
    const unsigned bb = __float_as_uint(b);
    if (b > 0.0f && (bb & 0x007fffffu) == 0u && (fabsf(a) >= 1e-30f) && fabsf(a) < 1e30f)
        return floorf(__fmul_rn(a, __frcp_rn(b)));
    return floordiv_f32_generic(a, b);
}

__device__ __forceinline__ float floordiv_f32_generic(float a, float b)
{
    float mod = fmodf(a, b);
    float div = __fdiv_rn(__fsub_rn(a, mod), b);
    if (mod != 0.0f && ((b < 0.0f) != (mod < 0.0f))) div = __fsub_rn(div, 1.0f);
    float fd;
    if (div != 0.0f) {
        fd = floorf(div);
        if (__fsub_rn(div, fd) > 0.5f) fd = __fadd_rn(fd, 1.0f);
    } else {
        fd = copysignf(0.0f, __fdiv_rn(a, b));
    }
    return fd;
}

// Query-to-trajectory distance exactly as the oracle / torch evaluate it in float32:
// l2: fl(fl(dy*dy) + fl(dx*dx)); l1: fl(|dy| + |dx|), dy = gy - py (focus.py:132-135).
__device__ __forceinline__ float knn_dist(float gy, float gx, float py, float px, int l1)
{
    float dy = __fsub_rn(gy, py), dx = __fsub_rn(gx, px);
    return l1 ? __fadd_rn(fabsf(dy), fabsf(dx))
              : __fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx));
}

__device__ __forceinline__ bool lex_less(float d1, int j1, float d2, int j2)
{
    return d1 < d2 || (d1 == d2 && j1 < j2);
}

// The four bilinear corners of one warped event (event_image_converter.py:354-386).
struct Corners {
    int idx[4];      // linear pixel index, -1 when the corner is out of bounds
    float fy, fx;    // fractional parts (may be slightly negative)
};

__device__ __forceinline__ Corners vote_corners(float wy, float wx, int H, int W)
{
    Corners c;
    float y1 = floorf(__fadd_rn(wy, kVoteEps));
    float x1 = floorf(__fadd_rn(wx, kVoteEps));
    c.fy = __fsub_rn(wy, y1);
    c.fx = __fsub_rn(wx, x1);
    // range tests in float so huge / NaN coordinates never reach the int conversion
    bool y0ok = (y1 >= 0.0f) && (y1 < (float)H);
    bool y1ok = (y1 >= -1.0f) && (y1 < (float)(H - 1));
    bool x0ok = (x1 >= 0.0f) && (x1 < (float)W);
    bool x1ok = (x1 >= -1.0f) && (x1 < (float)(W - 1));
    int iy = (y0ok || y1ok) ? (int)y1 : 0;
    int ix = (x0ok || x1ok) ? (int)x1 : 0;
    c.idx[0] = (y0ok && x0ok) ? iy * W + ix : -1;             // (y1,   x1)
    c.idx[1] = (y1ok && x0ok) ? (iy + 1) * W + ix : -1;       // (y1+1, x1)
    c.idx[2] = (y0ok && x1ok) ? iy * W + ix + 1 : -1;         // (y1,   x1+1)
    c.idx[3] = (y1ok && x1ok) ? (iy + 1) * W + ix + 1 : -1;   // (y1+1, x1+1)
    return c;
}

// The imager's outer padding (event_image_converter.py:339-380): the integer corner is shifted by
// (ph, pw) AFTER the floor, the fractional parts are those of the unshifted coordinate, and the
// bounds are those of the padded image (H, W).  Shifting the float coordinate instead would round.
__device__ __forceinline__ Corners vote_corners_padded(float wy, float wx, int H, int W, int ph, int pw)
{
    Corners c;
    const float yf = floorf(__fadd_rn(wy, kVoteEps)), xf = floorf(__fadd_rn(wx, kVoteEps));
    c.fy = __fsub_rn(wy, yf);
    c.fx = __fsub_rn(wx, xf);
    // |floor| < 2^23 keeps the shifted corner an exact float; anything larger is far outside
    const bool sane = fabsf(yf) < 8388608.0f && fabsf(xf) < 8388608.0f;
    const float y1 = yf + (float)ph, x1 = xf + (float)pw;
    bool y0ok = sane && (y1 >= 0.0f) && (y1 < (float)H);
    bool y1ok = sane && (y1 >= -1.0f) && (y1 < (float)(H - 1));
    bool x0ok = sane && (x1 >= 0.0f) && (x1 < (float)W);
    bool x1ok = sane && (x1 >= -1.0f) && (x1 < (float)(W - 1));
    int iy = (y0ok || y1ok) ? (int)y1 : 0;
    int ix = (x0ok || x1ok) ? (int)x1 : 0;
    c.idx[0] = (y0ok && x0ok) ? iy * W + ix : -1;
    c.idx[1] = (y1ok && x0ok) ? (iy + 1) * W + ix : -1;
    c.idx[2] = (y0ok && x1ok) ? iy * W + ix + 1 : -1;
    c.idx[3] = (y1ok && x1ok) ? (iy + 1) * W + ix + 1 : -1;
    return c;
}

// ---- per-event helpers shared by event_stage.cu and tile_stage.cu ----
struct EventRow { float y, x, t, p, bin, valid; };

__device__ __forceinline__ EventRow load_event(const float *row)
{
    float2 a = ld_stream_f2(row), b = ld_stream_f2(row + 2), c = ld_stream_f2(row + 4);
    EventRow e{a.x, a.y, b.x, b.y, c.x, c.y};
    return e;
}

// LUT cell of an event (focus.py:185-187). Returns false when outside the table.
__device__ __forceinline__ bool lut_cell(const EventRow &e, const Geom &g, int64_t b, int64_t *cell)
{
    float fs = (float)g.s;
    float fy = floordiv_f32(e.y, fs), fx = floordiv_f32(e.x, fs);
    float ft = truncf(e.bin);
    if (!(ft >= 0.0f && ft < (float)g.nb && fy >= 0.0f && fy < (float)g.Hq && fx >= 0.0f &&
          fx < (float)g.Wq))
        return false;
    *cell = ((b * g.nb + (int)ft) * g.Hq + (int)fy) * g.Wq + (int)fx;
    return true;
}

// weight of a warped event (focus.py:201-214), no gradient flows through it
__device__ __forceinline__ float event_weight(const EventRow &e, float wy, float wx, float tref,
                                              const Geom &g)
{
    float w = e.valid;
    if (g.scale_dt) {
        float dt = fminf(fmaxf(fabsf(__fsub_rn(e.t, tref)), 0.0f), 1.0f);
        w = __fmul_rn(__fsub_rn(1.0f, dt), w);
    }
    if (g.mask_border) {
        if (wy > (float)g.H || wx > (float)g.W || wy < 0.0f || wx < 0.0f) w = 0.0f;
    }
    return w;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum (fixed tree order -> deterministic); result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double *smem32)
{
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    v = warp_sum(v);
    if (lane == 0) smem32[wid] = v;
    __syncthreads();
    int nw = (blockDim.x * blockDim.y + 31) >> 5;
    double r = 0.0;
    if (wid == 0) {
        r = lane < nw ? smem32[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}
#endif  // __CUDACC__

}  // namespace cmax
