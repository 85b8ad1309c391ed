// api.cu - the extern "C" boundary declared in include/cmax_b200.h, geometry / workspace
// layout, the fused coeff_grid -> trajectories front end and the atomic micro-benchmark.
#include <string.h>

#include "cmax_common.cuh"

namespace cmax {

int launch_splat(int mode, const float *events, const float *weight, int64_t nb, int64_t M,
                 int64_t stride, int H, int W, int ph, int pw, float *out, long long *out_i64,
                 cudaStream_t st);
int launch_blur(const float *raw, float *out, int64_t planes, int H, int W, float sigma,
                cudaStream_t st);
extern const int *g_last_work_count;

// ---------------------------------------------------------------------------------------------
// stage timing / launch counting (single-threaded caller per process, like the reference)
// ---------------------------------------------------------------------------------------------
static const char *kStageNames[ST_COUNT] = {
    "bin_points", "knn_select", "event_forward", "image_forward", "smooth_forward", "finalize",
    "image_backward", "smooth_backward", "event_backward", "lut_backward", "traj_forward",
    "traj_backward", "pack_events"};
constexpr int kMaxTimed = 8192;
static bool g_timing = false;
static cudaEvent_t g_ev[kMaxTimed][2];
static int g_ev_stage[kMaxTimed];
static int g_ev_created = 0, g_ev_used = 0;
static int g_open[ST_COUNT];
static long long g_launches = 0;

void count_launch(int n) { g_launches += n; }

void stage_begin(int stage, cudaStream_t st)
{
    g_open[stage] = -1;
    if (!g_timing || g_ev_used >= kMaxTimed) return;
    if (g_ev_used >= g_ev_created) {
        if (cudaEventCreate(&g_ev[g_ev_created][0]) != cudaSuccess) return;
        if (cudaEventCreate(&g_ev[g_ev_created][1]) != cudaSuccess) return;
        ++g_ev_created;
    }
    const int i = g_ev_used++;
    g_ev_stage[i] = stage;
    g_open[stage] = i;
    cudaEventRecord(g_ev[i][0], st);
}

void stage_end(int stage, cudaStream_t st)
{
    if (g_open[stage] >= 0) cudaEventRecord(g_ev[g_open[stage]][1], st);
    g_open[stage] = -1;
}

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
void knn_geom(int H, int W, int s, int64_t n, int K, Geom *g)
{
    g->H = H;
    g->W = W;
    g->s = s;
    g->K = K;
    g->n = n;
    g->Hq = (H + s - 1) / s;                    // len(arange(0, H, s)), focus.py:118-124
    g->Wq = (W + s - 1) / s;
    g->q = g->Hq * g->Wq;
    g->off = (float)s / 2.0f - 0.5f;
    // cell edge: the smallest multiple of s with at most kMaxCells cells and, for sparse point
    // sets, roughly one point per cell or more
    int m = 1;
    while (true) {
        int cs = s * m;
        int64_t nc = (int64_t)((H + cs - 1) / cs) * ((W + cs - 1) / cs);
        double per_cell = (double)n / (double)nc;
        if (nc <= kMaxCells && (per_cell >= 0.75 || nc <= 64)) break;
        ++m;
    }
    g->cs = (float)(s * m);
    g->inv_cs = 1.0f / g->cs;
    g->Hc = (H + s * m - 1) / (s * m);
    g->Wc = (W + s * m - 1) / (s * m);
    g->NC = g->Hc * g->Wc;
    // initial window: radius that holds K points at the mean density
    double dens = (double)n / ((double)H * W);
    double rk = sqrt((double)K / (3.14159265358979 * (dens > 0 ? dens : 1e-9)));
    int r0 = (int)ceil((rk - 0.5 * g->cs) / g->cs);
    g->r0 = r0 < 0 ? 0 : (r0 > 8 ? 8 : r0);
    g->r_fast = g->r0 + 2;
}

int make_geom(const CmaxConfig *c, int64_t B, int64_t M, int64_t n, int64_t npos, Geom *g)
{
    if (!c) return CMAX_ERR_BAD_CONFIG;
    if (c->height < 3 || c->width < 3 || c->num_tref < 1 || c->num_bins < 1 || c->num_knn < 1 ||
        c->lut_superpixel_size < 1)
        return CMAX_ERR_BAD_CONFIG;
    if ((unsigned)c->focus_loss_norm > 1u || (unsigned)c->dist_norm > 1u ||
        (unsigned)c->interpolation_scheme > 1u || (unsigned)c->smooth_type > 1u ||
        (unsigned)c->focus_functional > 1u)
        return CMAX_ERR_BAD_CONFIG;
    // focus.py:49-51
    if (c->num_tref != 1 && (c->scale_iwe_by_dt || c->polarity_aware_batching ||
                             c->smooth_type == CMAX_SMOOTH_ON_FLOW_TO_NEXT))
        return CMAX_ERR_BAD_CONFIG;
    // focus.py:170 builds flow_to_next only for smooth_weight > 0, and calculate_smooth_loss (:232-246)
    // then dereferences None for a negative weight: the reference crashes on this combination
    if (c->smooth_type == CMAX_SMOOTH_ON_FLOW_TO_NEXT && c->smooth_weight < 0.0f) return CMAX_ERR_BAD_CONFIG;
    if (B < 1 || M < 0 || n < 1) return CMAX_ERR_BAD_SHAPE;
    if (c->num_knn > n) return CMAX_ERR_BAD_SHAPE;
    if (c->polarity_aware_batching && (npos < 0 || npos > M)) return CMAX_ERR_BAD_SHAPE;
    if (c->num_knn > kMaxKnn || c->num_tref > kMaxTref) return CMAX_ERR_UNSUPPORTED;
    if (B > 65535 || c->num_bins > 65535 || n > (int64_t)INT32_MAX / 2 || B * c->num_bins > (int64_t)INT32_MAX)
        return CMAX_ERR_UNSUPPORTED;
    memset(g, 0, sizeof(*g));
    knn_geom(c->height, c->width, c->lut_superpixel_size, n, c->num_knn, g);
    g->R = c->num_tref;
    g->nb = c->num_bins;
    g->P = c->polarity_aware_batching ? 2 : 1;
    g->l1dist = c->dist_norm == CMAX_NORM_L1;
    g->l2focus = c->focus_loss_norm == CMAX_NORM_L2;
    g->scale_dt = c->scale_iwe_by_dt != 0;
    g->mask_border = c->mask_image_border != 0;
    g->pab = c->polarity_aware_batching != 0;
    g->iwd = c->interpolation_scheme == CMAX_INTERP_IWD && c->num_knn > 1;    // focus.py:145-163
    g->smooth_next = c->smooth_type == CMAX_SMOOTH_ON_FLOW_TO_NEXT;
    g->det = c->deterministic != 0;
    g->variance = c->focus_functional == CMAX_FOCUS_VARIANCE;
    g->fuse_image = (c->backward_follows != 0) && !g->variance;
    g->smooth_w = c->smooth_weight;
    g->B = B;
    g->M = M;
    g->S = B * c->num_bins;
    g->npos = npos;
    // packed (tile-binned) event layout: source tiles of ct x ct LUT cells, ~32 px on a side
    g->ct = 32 / g->s > 0 ? 32 / g->s : 1;
    g->nty = (g->Hq + g->ct - 1) / g->ct;
    g->ntx = (g->Wq + g->ct - 1) / g->ct;
    g->nt = g->nty * g->ntx;
    return CMAX_OK;
}

template <class Take>
static void take_knn(const Geom &g, Layout &L, Take &take)
{
    const size_t tiles = (size_t)((g.Wq + kKnnTileW - 1) / kKnnTileW) * ((g.Hq + kKnnTileH - 1) / kKnnTileH);
    L.cell_start = take(sizeof(int) * g.S * (g.NC + 1));
    L.sorted = take(sizeof(float4) * g.S * g.n);
    L.recs = take(sizeof(float4) * g.S * g.n);
    L.sorted_j = take(sizeof(int) * g.S * g.n);
    L.tau = take(sizeof(float) * g.S * g.q);
    L.jcut = take(sizeof(int) * g.S * g.q);
    L.wsum = take(g.iwd ? sizeof(float) * g.S * g.q : 16);
    L.tau_max = take(sizeof(unsigned) * g.S);
    L.tile_max = take(sizeof(unsigned) * g.S * tiles);
    L.worklist = take(sizeof(int) * g.S * g.q);
    L.worklist2 = take(sizeof(int) * g.S * g.q);
    L.work_count = take(sizeof(int) * 16);        // [0] work-list length, [1..7] miss reasons, [8] second list
}

Layout make_knn_layout(const Geom &g)
{
    Layout L;
    memset(&L, 0, sizeof(L));
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes);
        return o;
    };
    take_knn(g, L, take);
    L.total = off;
    return L;
}

Layout make_layout(const Geom &g)
{
    Layout L;
    memset(&L, 0, sizeof(L));
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes);
        return o;
    };
    const int64_t planes = g.B * g.R * g.P;
    const int64_t npix = planes * (int64_t)g.H * g.W;
    const int64_t nlut = g.S * g.q * g.R * 2;
    const int64_t nf2n = g.smooth_next ? g.B * (g.nb - 1) * g.q * 2 : 0;
    L.n_img_blocks = (int)(planes * ((g.H + kImgTile - 1) / kImgTile) * ((g.W + kImgTile - 1) / kImgTile));
    const int64_t sm_imgs = g.smooth_next ? g.B * (g.nb - 1) : g.S * g.R;
    L.n_sm_blocks = (int)(sm_imgs * ((g.Wq + 31) / 32) * ((g.Hq + 7) / 8));
    L.header = take(1024);
    L.focus_partials = take(sizeof(double) * 2 * L.n_img_blocks);   // [sum | sum of squares]
    L.plane_stats = take(sizeof(double) * 2 * planes);              // per plane: mean, variance
    L.smooth_partials = take(sizeof(double) * (L.n_sm_blocks > 0 ? L.n_sm_blocks : 1));
    take_knn(g, L, take);
    L.bpart = take(sizeof(float2) * g.S * g.n * (g.R + (g.smooth_next ? 1 : 0)));
    L.lut = take(sizeof(float) * nlut);
    L.f2n = take(sizeof(float) * (nf2n > 0 ? nf2n : 4));
    L.raw = take(sizeof(float) * npix);
    L.raw_i64 = take(g.det ? sizeof(long long) * npix : 16);
    L.dimg = take(sizeof(float) * npix);
    L.dlut = take(sizeof(float) * nlut);
    L.dlut_i64 = take(g.det ? sizeof(long long) * nlut : 16);
    L.df2n = take(sizeof(float) * (nf2n > 0 ? nf2n : 4));
    L.total = off;
    return L;
}

// ---------------------------------------------------------------------------------------------
// front end: coeff_grid -> trajectories (trajectory_net.py:101-119) and its adjoint
// ---------------------------------------------------------------------------------------------
// one thread per (b, tile j): K coefficients per axis in registers, basis table phi[n_t, K]
// (already phi(t) - phi(anchor)) broadcast from shared memory.
constexpr int kMaxBasis = 16;

template <int KB>       // register budget for the coefficients: 4 (polynomial K <= 4) or kMaxBasis
__global__ void __launch_bounds__(128)
traj_forward_kernel(const float *__restrict__ cg, const float *__restrict__ phi, int64_t S, int K,
                    int H, int W, int patch, int ny, int nx, int n_t, int xy, int add_off,
                    float *__restrict__ out)
{
    extern __shared__ float s_phi[];
    for (int i = threadIdx.x; i < n_t * K; i += blockDim.x) s_phi[i] = phi[i];
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = blockIdx.y;
    const int64_t n = (int64_t)ny * nx;
    if (j >= n) return;
    const int ty = (int)(j / nx), tx = (int)(j - (int64_t)ty * nx);
    const int py = patch / 2 + ty * patch, px = patch / 2 + tx * patch;     // trajectories.py:8-13
    float cy[KB], cx[KB];
    const int64_t HW = (int64_t)H * W;
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        if (k < K) {
            float a0 = 0.0f, a1 = 0.0f;
            for (int64_t sc = 0; sc < S; ++sc) {                        // sum over scales (basis.py:46)
                const float *base = cg + ((b * S + sc) * 2 * K) * HW + (int64_t)py * W + px;
                a0 += __ldg(base + (int64_t)k * HW);
                a1 += __ldg(base + (int64_t)(K + k) * HW);
            }
            cy[k] = xy ? a1 : a0;
            cx[k] = xy ? a0 : a1;
        }
    }
    float2 *o = reinterpret_cast<float2 *>(out) + b * (int64_t)n_t * n + j;
    for (int t = 0; t < n_t; ++t) {
        float y = 0.0f, x = 0.0f;
#pragma unroll
        for (int k = 0; k < KB; ++k)
            if (k < K) {
                y += s_phi[t * K + k] * cy[k];
                x += s_phi[t * K + k] * cx[k];
            }
        if (add_off) { y += (float)py; x += (float)px; }
        o[(int64_t)t * n] = make_float2(y, x);
    }
}

template <int KB>
__global__ void __launch_bounds__(128)
traj_backward_kernel(const float *__restrict__ dtraj, const float *__restrict__ phi, int64_t S,
                     int K, int H, int W, int patch, int ny, int nx, int n_t, int xy,
                     float *__restrict__ dcg)
{
    extern __shared__ float s_phi[];
    for (int i = threadIdx.x; i < n_t * K; i += blockDim.x) s_phi[i] = phi[i];
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = blockIdx.y;
    const int64_t n = (int64_t)ny * nx;
    if (j >= n) return;
    const int ty = (int)(j / nx), tx = (int)(j - (int64_t)ty * nx);
    const int py = patch / 2 + ty * patch, px = patch / 2 + tx * patch;
    float gy[KB], gx[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) gy[k] = gx[k] = 0.0f;
    const float2 *d = reinterpret_cast<const float2 *>(dtraj) + b * (int64_t)n_t * n + j;
    for (int t = 0; t < n_t; ++t) {
        const float2 v = __ldg(d + (int64_t)t * n);
#pragma unroll
        for (int k = 0; k < KB; ++k)
            if (k < K) {
                gy[k] += s_phi[t * K + k] * v.x;
                gx[k] += s_phi[t * K + k] * v.y;
            }
    }
    const int64_t HW = (int64_t)H * W;
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        if (k < K) {
            for (int64_t sc = 0; sc < S; ++sc) {
                float *base = dcg + ((b * S + sc) * 2 * K) * HW + (int64_t)py * W + px;
                base[(int64_t)k * HW] = xy ? gx[k] : gy[k];
                base[(int64_t)(K + k) * HW] = xy ? gy[k] : gx[k];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// atomic micro-benchmark
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned hash32(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int MODE>
__global__ void __launch_bounds__(256)
atomic_bench_kernel(float *region, unsigned region_floats, int64_t n_ops)
{
    __shared__ float s_tile[MODE == 1 ? 12288 : 1];
    if (MODE == 1) {
        for (int i = threadIdx.x; i < 12288; i += blockDim.x) s_tile[i] = 0.0f;
        __syncthreads();
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < n_ops; i += stride) {
        unsigned h = hash32((unsigned)i * 2654435761u + 12345u);
        if (MODE == 3 || MODE == 4) {
            // vector reds as event_forward issues them: one request per image row of the vote
            // (MODE 3: the two x-adjacent corners as red.v2 on an 8-byte aligned pair; MODE 4: the
            // pair in the middle of a 16-byte aligned quad as red.v4 with zeros in the outer lanes)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                unsigned a = (h % (region_floats - 644u)) + k * 640u;
                if (MODE == 3) red_add_f32x2(region + (a & ~1u), 1.0f, 1.0f);
                else red_add_f32x4(region + (a & ~3u), 0.0f, 1.0f, 1.0f, 0.0f);
            }
            continue;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // mimic a bilinear vote: (p, p+1, p+W, p+W+1) with W = 640
            unsigned a = (h % (region_floats - 642u)) + (k & 1) + (k >> 1) * 640u;
            if (MODE == 0) atomicAdd(region + a, 1.0f);
            if (MODE == 1) atomicAdd(&s_tile[a % 12288u], 1.0f);
            if (MODE == 2)
                atomicAdd(reinterpret_cast<unsigned long long *>(region) + (a >> 1), 1ull);
        }
    }
    if (MODE == 1) {
        __syncthreads();
        for (int i = threadIdx.x; i < 12288; i += blockDim.x)
            if (s_tile[i] != 0.0f) atomicAdd(region + (i % region_floats), s_tile[i]);
    }
}

}  // namespace cmax

// ---------------------------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------------------------
using namespace cmax;

extern "C" {

int cmax_abi_version(void) { return CMAX_ABI_VERSION; }

const char *cmax_error_string(int code)
{
    switch (code) {
    case CMAX_OK: return "ok";
    case CMAX_ERR_BAD_CONFIG: return "invalid CmaxConfig (range, or a combination upstream focus.py:49-51 forbids)";
    case CMAX_ERR_BAD_SHAPE: return "inconsistent B / M / n / num_pos_events / num_knn";
    case CMAX_ERR_WORKSPACE: return "workspace missing, misaligned or too small";
    case CMAX_ERR_CUDA: return "CUDA launch failure";
    case CMAX_ERR_UNSUPPORTED: return "size not supported by this build (num_knn <= 192, num_tref <= 16, B <= 65535)";
    default: return "unknown error";
    }
}

size_t cmax_workspace_bytes(const CmaxConfig *cfg, int64_t B, int64_t M, int64_t n)
{
    Geom g;
    if (make_geom(cfg, B, M, n, 0, &g) != CMAX_OK) return 0;
    return make_layout(g).total;
}

int cmax_forward(const CmaxConfig *cfg, const float *trajectories, const float *times,
                 const float *events, int64_t B, int64_t M, int64_t n, int64_t num_pos_events,
                 float *iwes_out, float *losses_out, float *flow_lut_out, void *workspace,
                 size_t workspace_bytes, void *stream)
{
    DeviceGuard dev_guard(workspace);
    Geom g;
    int rc = make_geom(cfg, B, M, n, num_pos_events, &g);
    if (rc != CMAX_OK) return rc;
    if (!trajectories || !times || (!events && M > 0) || !iwes_out || !losses_out)
        return CMAX_ERR_BAD_SHAPE;
    const Layout L = make_layout(g);
    if (!workspace || ((uintptr_t)workspace & 255u) || workspace_bytes < L.total)
        return CMAX_ERR_WORKSPACE;
    char *ws = static_cast<char *>(workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(ws + L.header, 0, 1024, st);
    if ((rc = launch_lut_forward(g, L, trajectories, ws, flow_lut_out, nullptr, nullptr, st))) return rc;
    if ((rc = launch_event_forward(g, L, events, times, ws, st))) return rc;
    if ((rc = g.fuse_image ? launch_image_forward_backward(g, L, ws, iwes_out, st)
                           : launch_image_forward(g, L, ws, iwes_out, st))) return rc;
    if ((rc = launch_smooth_forward(g, L, ws, st))) return rc;
    return launch_finalize_losses(g, L, ws, losses_out, st);
}

int cmax_backward(const CmaxConfig *cfg, const float *trajectories, const float *times,
                  const float *events, int64_t B, int64_t M, int64_t n, int64_t num_pos_events,
                  const float *grad_loss, float *dtraj_out, void *workspace,
                  size_t workspace_bytes, void *stream)
{
    DeviceGuard dev_guard(workspace);
    Geom g;
    int rc = make_geom(cfg, B, M, n, num_pos_events, &g);
    if (rc != CMAX_OK) return rc;
    if (!trajectories || !times || (!events && M > 0) || !grad_loss || !dtraj_out)
        return CMAX_ERR_BAD_SHAPE;
    const Layout L = make_layout(g);
    if (!workspace || ((uintptr_t)workspace & 255u) || workspace_bytes < L.total)
        return CMAX_ERR_WORKSPACE;
    char *ws = static_cast<char *>(workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!g.fuse_image && (rc = launch_image_backward(g, L, ws, st))) return rc;   // else done by the forward
    if ((rc = launch_smooth_backward(g, L, grad_loss, ws, st))) return rc;
    if ((rc = launch_event_backward(g, L, events, times, grad_loss, ws, st))) return rc;
    return launch_lut_backward(g, L, trajectories, ws, dtraj_out, st);
}

// ---- phased entry points (event-sharded single-window mode) ------------------------
int cmax_workspace_section(const CmaxConfig *cfg, int64_t B, int64_t M, int64_t n, int32_t which,
                           size_t *offset_out, size_t *bytes_out, int32_t *is_int64_out)
{
    Geom g;
    int rc = make_geom(cfg, B, M, n, 0, &g);
    if (rc != CMAX_OK) return rc;
    if (!offset_out || !bytes_out || !is_int64_out) return CMAX_ERR_BAD_SHAPE;
    const Layout L = make_layout(g);
    const size_t npix = (size_t)(g.B * g.R * g.P) * g.H * g.W, nlut = (size_t)(g.S * g.q * g.R * 2);
    *is_int64_out = g.det ? 1 : 0;
    if (which == CMAX_SECTION_RAW_IWE) {
        *offset_out = g.det ? L.raw_i64 : L.raw;
        *bytes_out = npix * (g.det ? 8 : 4);
    } else if (which == CMAX_SECTION_DLUT) {
        *offset_out = g.det ? L.dlut_i64 : L.dlut;
        *bytes_out = nlut * (g.det ? 8 : 4);
    } else {
        return CMAX_ERR_BAD_CONFIG;
    }
    return CMAX_OK;
}

int cmax_forward_accumulate(const CmaxConfig *cfg, const float *trajectories, const float *times,
                            const float *events, int64_t B, int64_t M, int64_t n,
                            int64_t num_pos_events, float *flow_lut_out, void *workspace,
                            size_t workspace_bytes, void *stream)
{
    DeviceGuard dev_guard(workspace);
    Geom g;
    int rc = make_geom(cfg, B, M, n, num_pos_events, &g);
    if (rc != CMAX_OK) return rc;
    if (!trajectories || !times || (!events && M > 0)) return CMAX_ERR_BAD_SHAPE;
    const Layout L = make_layout(g);
    if (!workspace || ((uintptr_t)workspace & 255u) || workspace_bytes < L.total)
        return CMAX_ERR_WORKSPACE;
    char *ws = static_cast<char *>(workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(ws + L.header, 0, 1024, st);
    if ((rc = launch_lut_forward(g, L, trajectories, ws, flow_lut_out, nullptr, nullptr, st))) return rc;
    return launch_event_forward(g, L, events, times, ws, st, 1);
}

int cmax_forward_finish(const CmaxConfig *cfg, int64_t B, int64_t M, int64_t n, float *iwes_out,
                        float *losses_out, void *workspace, size_t workspace_bytes, void *stream)
{
    DeviceGuard dev_guard(workspace);
    Geom g;
    int rc = make_geom(cfg, B, M, n, 0, &g);
    if (rc != CMAX_OK) return rc;
    if (!iwes_out || !losses_out) return CMAX_ERR_BAD_SHAPE;
    const Layout L = make_layout(g);
    if (!workspace || ((uintptr_t)workspace & 255u) || workspace_bytes < L.total)
        return CMAX_ERR_WORKSPACE;
    char *ws = static_cast<char *>(workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((rc = launch_event_forward(g, L, nullptr, nullptr, ws, st, 2))) return rc;
    if ((rc = launch_image_forward(g, L, ws, iwes_out, st))) return rc;
    if ((rc = launch_smooth_forward(g, L, ws, st))) return rc;
    return launch_finalize_losses(g, L, ws, losses_out, st);
}

int cmax_backward_accumulate(const CmaxConfig *cfg, const float *trajectories, const float *times,
                             const float *events, int64_t B, int64_t M, int64_t n,
                             int64_t num_pos_events, const float *grad_loss, int32_t include_smooth,
                             void *workspace, size_t workspace_bytes, void *stream)
{
    DeviceGuard dev_guard(workspace);
    Geom g;
    int rc = make_geom(cfg, B, M, n, num_pos_events, &g);
    if (rc != CMAX_OK) return rc;
    if (!trajectories || !times || (!events && M > 0) || !grad_loss) return CMAX_ERR_BAD_SHAPE;
    const Layout L = make_layout(g);
    if (!workspace || ((uintptr_t)workspace & 255u) || workspace_bytes < L.total)
        return CMAX_ERR_WORKSPACE;
    char *ws = static_cast<char *>(workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((rc = launch_image_backward(g, L, ws, st))) return rc;
    if ((rc = launch_smooth_backward(g, L, grad_loss, ws, st))) return rc;
    // float mode: the smoothness gradient sits in dLUT and must enter the sum over ranks once
    if (!include_smooth && !g.det)
        cudaMemsetAsync(ws + L.dlut, 0, sizeof(float) * g.S * g.q * g.R * 2, st);
    return launch_event_backward(g, L, events, times, grad_loss, ws, st, 1);
}

int cmax_backward_finish(const CmaxConfig *cfg, const float *trajectories, int64_t B, int64_t M,
                         int64_t n, const float *grad_loss, float *dtraj_out, void *workspace,
                         size_t workspace_bytes, void *stream)
{
    DeviceGuard dev_guard(workspace);
    Geom g;
    int rc = make_geom(cfg, B, M, n, 0, &g);
    if (rc != CMAX_OK) return rc;
    if (!trajectories || !grad_loss || !dtraj_out) return CMAX_ERR_BAD_SHAPE;
    const Layout L = make_layout(g);
    if (!workspace || ((uintptr_t)workspace & 255u) || workspace_bytes < L.total)
        return CMAX_ERR_WORKSPACE;
    char *ws = static_cast<char *>(workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((rc = launch_event_backward(g, L, nullptr, nullptr, grad_loss, ws, st, 2))) return rc;
    return launch_lut_backward(g, L, trajectories, ws, dtraj_out, st);
}

static int pack_supported(const Geom &g)
{
    if (g.nb > 256 || g.Hq > 4096 || g.Wq > 4096 || g.M > (int64_t)INT32_MAX) return CMAX_ERR_UNSUPPORTED;
    if ((int64_t)g.P * g.nt > 12288) return CMAX_ERR_UNSUPPORTED;     // segment counters live in 48 KB of smem
    if (g.S * (int64_t)g.q * g.R >= (int64_t)INT32_MAX / 2) return CMAX_ERR_UNSUPPORTED;   // 32-bit LUT indices
    return CMAX_OK;
}

int cmax_pack_layout(const CmaxConfig *cfg, int32_t out_host[4])
{
    Geom g;
    int rc = make_geom(cfg, 1, 0, cfg ? cfg->num_knn : 1, 0, &g);
    if (rc != CMAX_OK) return rc;
    if (!out_host) return CMAX_ERR_BAD_SHAPE;
    if ((rc = pack_supported(g))) return rc;
    out_host[0] = g.ct;
    out_host[1] = g.nty;
    out_host[2] = g.ntx;
    out_host[3] = g.P;
    return CMAX_OK;
}

int cmax_pack_events(const CmaxConfig *cfg, const float *events, int64_t B, int64_t M,
                     int64_t num_pos_events, float *records_out, int32_t *seg_start_out,
                     int32_t *scratch, int64_t *skipped_out, void *stream)
{
    DeviceGuard dev_guard(seg_start_out);
    Geom g;
    int rc = make_geom(cfg, B, M, cfg ? cfg->num_knn : 1, num_pos_events, &g);
    if (rc != CMAX_OK) return rc;
    if ((rc = pack_supported(g))) return rc;
    if ((!events && M > 0) || (!records_out && M > 0) || !seg_start_out || !scratch) return CMAX_ERR_BAD_SHAPE;
    if (((uintptr_t)records_out & 15u)) return CMAX_ERR_WORKSPACE;
    return launch_pack_events(g, events, reinterpret_cast<float4 *>(records_out), seg_start_out, scratch,
                              reinterpret_cast<long long *>(skipped_out), static_cast<cudaStream_t>(stream));
}

int64_t cmax_expand_scratch_ints(const CmaxConfig *cfg, int64_t B, int64_t records_stride)
{
    Geom g;
    if (make_geom(cfg, B, records_stride, cfg ? cfg->num_knn : 1, 0, &g) != CMAX_OK) return 0;
    const int64_t span = records_stride > g.P * g.nt + 1 ? records_stride : g.P * g.nt + 1;
    return B * ((span + 255) / 256);
}

int cmax_expand_compact(const CmaxConfig *cfg, const float *coords, const int32_t *fine_start,
                        const int64_t *sample_off, int64_t B, int64_t records_stride,
                        float *records_out, int32_t *seg_start_out, int32_t *scratch, void *stream)
{
    DeviceGuard dev_guard(seg_start_out);
    Geom g;
    int rc = make_geom(cfg, B, records_stride, cfg ? cfg->num_knn : 1, 0, &g);
    if (rc != CMAX_OK) return rc;
    if ((rc = pack_supported(g))) return rc;
    if (!fine_start || !sample_off || !seg_start_out || !scratch || (records_stride > 0 && (!coords || !records_out)))
        return CMAX_ERR_BAD_SHAPE;
    if (((uintptr_t)records_out & 15u)) return CMAX_ERR_WORKSPACE;
    return launch_expand_compact(g, coords, fine_start, reinterpret_cast<const long long *>(sample_off),
                                 records_stride, reinterpret_cast<float4 *>(records_out), seg_start_out, scratch,
                                 static_cast<cudaStream_t>(stream));
}

int cmax_expand_bitpacked(const CmaxConfig *cfg, const uint32_t *words, const int32_t *fine_start,
                          const uint32_t *run_hdr, const int32_t *run_word, const int64_t *word_off, int64_t B,
                          int64_t records_stride, float *records_out, int32_t *seg_start_out, int32_t *scratch,
                          void *stream)
{
    DeviceGuard dev_guard(seg_start_out);
    Geom g;
    int rc = make_geom(cfg, B, records_stride, cfg ? cfg->num_knn : 1, 0, &g);
    if (rc != CMAX_OK) return rc;
    if ((rc = pack_supported(g))) return rc;
    if (!fine_start || !run_hdr || !run_word || !word_off || !seg_start_out || !scratch ||
        (records_stride > 0 && (!words || !records_out)))
        return CMAX_ERR_BAD_SHAPE;
    if (((uintptr_t)records_out & 15u) || ((uintptr_t)run_hdr & 15u)) return CMAX_ERR_WORKSPACE;
    return launch_expand_bitpacked(g, words, fine_start, run_hdr, run_word,
                                   reinterpret_cast<const long long *>(word_off), records_stride,
                                   reinterpret_cast<float4 *>(records_out), seg_start_out, scratch,
                                   static_cast<cudaStream_t>(stream));
}

int cmax_forward_packed(const CmaxConfig *cfg, const float *trajectories, const float *times,
                        const float *records, const int32_t *seg_start, int64_t B, int64_t M,
                        int64_t n, float *iwes_out, float *losses_out, float *flow_lut_out,
                        void *workspace, size_t workspace_bytes, void *stream)
{
    DeviceGuard dev_guard(workspace);
    Geom g;
    int rc = make_geom(cfg, B, M, n, 0, &g);
    if (rc != CMAX_OK) return rc;
    if ((rc = pack_supported(g))) return rc;
    if (!trajectories || !times || (!records && M > 0) || !seg_start || !iwes_out || !losses_out)
        return CMAX_ERR_BAD_SHAPE;
    if (((uintptr_t)records & 15u)) return CMAX_ERR_BAD_SHAPE;
    const Layout L = make_layout(g);
    if (!workspace || ((uintptr_t)workspace & 255u) || workspace_bytes < L.total)
        return CMAX_ERR_WORKSPACE;
    char *ws = static_cast<char *>(workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(ws + L.header, 0, 1024, st);
    if ((rc = launch_lut_forward(g, L, trajectories, ws, flow_lut_out, nullptr, nullptr, st))) return rc;
    if ((rc = launch_event_forward_packed(g, L, reinterpret_cast<const float4 *>(records), seg_start,
                                          times, ws, st))) return rc;
    if ((rc = g.fuse_image ? launch_image_forward_backward(g, L, ws, iwes_out, st)
                           : launch_image_forward(g, L, ws, iwes_out, st))) return rc;
    if ((rc = launch_smooth_forward(g, L, ws, st))) return rc;
    return launch_finalize_losses(g, L, ws, losses_out, st);
}

int cmax_backward_packed(const CmaxConfig *cfg, const float *trajectories, const float *times,
                         const float *records, const int32_t *seg_start, int64_t B, int64_t M,
                         int64_t n, const float *grad_loss, float *dtraj_out, void *workspace,
                         size_t workspace_bytes, void *stream)
{
    DeviceGuard dev_guard(workspace);
    Geom g;
    int rc = make_geom(cfg, B, M, n, 0, &g);
    if (rc != CMAX_OK) return rc;
    if ((rc = pack_supported(g))) return rc;
    if (!trajectories || !times || (!records && M > 0) || !seg_start || !grad_loss || !dtraj_out)
        return CMAX_ERR_BAD_SHAPE;
    const Layout L = make_layout(g);
    if (!workspace || ((uintptr_t)workspace & 255u) || workspace_bytes < L.total)
        return CMAX_ERR_WORKSPACE;
    char *ws = static_cast<char *>(workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!g.fuse_image && (rc = launch_image_backward(g, L, ws, st))) return rc;   // else done by the forward
    if ((rc = launch_smooth_backward(g, L, grad_loss, ws, st))) return rc;
    if ((rc = launch_event_backward_packed(g, L, reinterpret_cast<const float4 *>(records), seg_start,
                                           times, grad_loss, ws, st))) return rc;
    return launch_lut_backward(g, L, trajectories, ws, dtraj_out, st);
}

int cmax_create_iwe_padded(const float *events, const float *weight, int64_t nb, int64_t M,
                           int64_t row_stride, int32_t H, int32_t W, int32_t pad_y, int32_t pad_x,
                           float sigma, float *out, float *scratch, int64_t *scratch_i64,
                           int32_t deterministic, void *stream)
{
    DeviceGuard dev_guard(out);
    if (nb < 0 || M < 0 || row_stride < 2 || H < 1 || W < 1 || !out || (!events && M > 0))
        return CMAX_ERR_BAD_SHAPE;
    if (pad_y < 0 || pad_x < 0 || 2 * (int64_t)pad_y >= H || 2 * (int64_t)pad_x >= W)
        return CMAX_ERR_BAD_SHAPE;                       // H, W are the padded sizes
    if (sigma > 0.0f && (!scratch || H < 2 || W < 2)) return CMAX_ERR_WORKSPACE;
    if (deterministic && !scratch_i64) return CMAX_ERR_WORKSPACE;
    if (nb > 65535) return CMAX_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t count = nb * (int64_t)H * W;
    if (count == 0) return CMAX_OK;
    float *raw = sigma > 0.0f ? scratch : out;
    int rc;
    if (deterministic) {
        cudaMemsetAsync(scratch_i64, 0, sizeof(long long) * count, st);
        if ((rc = launch_splat(1, events, weight, nb, M, row_stride, H, W, pad_y, pad_x, nullptr,
                               reinterpret_cast<long long *>(scratch_i64), st))) return rc;
        if ((rc = launch_fix_to_float(reinterpret_cast<long long *>(scratch_i64), raw, count, st))) return rc;
    } else {
        cudaMemsetAsync(raw, 0, sizeof(float) * count, st);
        if ((rc = launch_splat(0, events, weight, nb, M, row_stride, H, W, pad_y, pad_x, raw, nullptr, st)))
            return rc;
    }
    if (sigma > 0.0f) return launch_blur(raw, out, nb, H, W, sigma, st);
    return check_launch();
}

int cmax_create_iwe(const float *events, const float *weight, int64_t nb, int64_t M,
                    int64_t row_stride, int32_t H, int32_t W, float sigma, float *out,
                    float *scratch, int64_t *scratch_i64, int32_t deterministic, void *stream)
{
    return cmax_create_iwe_padded(events, weight, nb, M, row_stride, H, W, 0, 0, sigma, out, scratch,
                                  scratch_i64, deterministic, stream);
}

int cmax_count_image_padded(const float *events, int64_t nb, int64_t M, int64_t row_stride, int32_t H,
                            int32_t W, int32_t pad_y, int32_t pad_x, int64_t *out, void *stream)
{
    DeviceGuard dev_guard(out);
    if (nb < 0 || M < 0 || row_stride < 2 || H < 1 || W < 1 || !out || (!events && M > 0))
        return CMAX_ERR_BAD_SHAPE;
    if (pad_y < 0 || pad_x < 0 || 2 * (int64_t)pad_y >= H || 2 * (int64_t)pad_x >= W)
        return CMAX_ERR_BAD_SHAPE;
    if (nb > 65535) return CMAX_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(out, 0, sizeof(long long) * nb * (int64_t)H * W, st);
    return launch_splat(2, events, nullptr, nb, M, row_stride, H, W, pad_y, pad_x, nullptr,
                        reinterpret_cast<long long *>(out), st);
}

int cmax_count_image(const float *events, int64_t nb, int64_t M, int64_t row_stride, int32_t H,
                     int32_t W, int64_t *out, void *stream)
{
    return cmax_count_image_padded(events, nb, M, row_stride, H, W, 0, 0, out, stream);
}

static int knn_only_geom(int32_t H, int32_t W, int32_t s, int64_t S, int64_t n, int32_t K, Geom *g)
{
    if (H < 1 || W < 1 || s < 1 || S < 1 || n < 1 || K < 1 || K > n) return CMAX_ERR_BAD_SHAPE;
    if (K > kMaxKnn || S > 65535) return CMAX_ERR_UNSUPPORTED;
    memset(g, 0, sizeof(*g));
    knn_geom(H, W, s, n, K, g);
    g->R = 0;
    g->nb = (int)S;
    g->B = 1;
    g->S = S;
    g->P = 1;
    return CMAX_OK;
}

size_t cmax_knn_workspace_bytes(int32_t H, int32_t W, int32_t s, int64_t S, int64_t n, int32_t K)
{
    Geom g;
    if (knn_only_geom(H, W, s, S, n, K, &g) != CMAX_OK) return 0;
    return make_knn_layout(g).total;
}

int cmax_knn_indices(const float *points, int64_t S, int64_t n, int32_t H, int32_t W, int32_t s,
                     int32_t K, int32_t dist_norm, int32_t *ind_out, float *dist_out,
                     void *workspace, size_t workspace_bytes, void *stream)
{
    DeviceGuard dev_guard(workspace);
    Geom g;
    int rc = knn_only_geom(H, W, s, S, n, K, &g);
    if (rc != CMAX_OK) return rc;
    if (!points || !ind_out) return CMAX_ERR_BAD_SHAPE;
    g.l1dist = dist_norm == CMAX_NORM_L1;
    const Layout L = make_knn_layout(g);
    if (!workspace || ((uintptr_t)workspace & 255u) || workspace_bytes < L.total)
        return CMAX_ERR_WORKSPACE;
    return launch_lut_forward(g, L, points, static_cast<char *>(workspace), nullptr, ind_out,
                              dist_out, static_cast<cudaStream_t>(stream));
}

static int traj_check(int64_t B, int64_t S, int32_t K, int32_t H, int32_t W, int32_t patch, int32_t n_t)
{
    if (B < 1 || S < 1 || K < 1 || H < 1 || W < 1 || patch < 1 || n_t < 1) return CMAX_ERR_BAD_SHAPE;
    if (patch / 2 >= H || patch / 2 >= W) return CMAX_ERR_BAD_SHAPE;
    if (K > kMaxBasis || B > 65535 || (size_t)n_t * K * 4 > 48 * 1024) return CMAX_ERR_UNSUPPORTED;
    return CMAX_OK;
}

int cmax_trajectories_forward(const float *coeff_grid, const float *phi, int64_t B, int64_t S,
                              int32_t K, int32_t H, int32_t W, int32_t patch, int32_t n_t,
                              int32_t xy_order, int32_t add_offsets, float *trajectories_out,
                              void *stream)
{
    DeviceGuard dev_guard(trajectories_out);
    int rc = traj_check(B, S, K, H, W, patch, n_t);
    if (rc) return rc;
    if (!coeff_grid || !phi || !trajectories_out) return CMAX_ERR_BAD_SHAPE;
    const int o = patch / 2;
    const int ny = (H - o + patch - 1) / patch, nx = (W - o + patch - 1) / patch;
    dim3 grid((unsigned)(((int64_t)ny * nx + 127) / 128), (unsigned)B);
    StageScope sc(ST_TRAJ_FWD, static_cast<cudaStream_t>(stream));
    count_launch();
    if (K <= 4)       // 8 instead of 32 coefficient registers: twice the resident warps for the common orders
        traj_forward_kernel<4><<<grid, 128, sizeof(float) * n_t * K, static_cast<cudaStream_t>(stream)>>>(
            coeff_grid, phi, S, K, H, W, patch, ny, nx, n_t, xy_order, add_offsets, trajectories_out);
    else
        traj_forward_kernel<kMaxBasis><<<grid, 128, sizeof(float) * n_t * K, static_cast<cudaStream_t>(stream)>>>(
            coeff_grid, phi, S, K, H, W, patch, ny, nx, n_t, xy_order, add_offsets, trajectories_out);
    return check_launch();
}

int cmax_trajectories_backward(const float *dtraj, const float *phi, int64_t B, int64_t S,
                               int32_t K, int32_t H, int32_t W, int32_t patch, int32_t n_t,
                               int32_t xy_order, float *dcoeff_grid_out, void *stream)
{
    DeviceGuard dev_guard(dcoeff_grid_out);
    int rc = traj_check(B, S, K, H, W, patch, n_t);
    if (rc) return rc;
    if (!dtraj || !phi || !dcoeff_grid_out) return CMAX_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int o = patch / 2;
    const int ny = (H - o + patch - 1) / patch, nx = (W - o + patch - 1) / patch;
    cudaMemsetAsync(dcoeff_grid_out, 0, sizeof(float) * B * S * 2 * K * (int64_t)H * W, st);
    dim3 grid((unsigned)(((int64_t)ny * nx + 127) / 128), (unsigned)B);
    StageScope sc(ST_TRAJ_BWD, st);
    count_launch();
    if (K <= 4)
        traj_backward_kernel<4><<<grid, 128, sizeof(float) * n_t * K, st>>>(
            dtraj, phi, S, K, H, W, patch, ny, nx, n_t, xy_order, dcoeff_grid_out);
    else
        traj_backward_kernel<kMaxBasis><<<grid, 128, sizeof(float) * n_t * K, st>>>(
            dtraj, phi, S, K, H, W, patch, ny, nx, n_t, xy_order, dcoeff_grid_out);
    return check_launch();
}

int cmax_atomic_microbench(float *region, int64_t region_floats, int64_t n_ops, int32_t mode,
                           void *stream)
{
    DeviceGuard dev_guard(region);
    if (!region || region_floats < 1024 || region_floats > 0x7fffffff || n_ops < 0)
        return CMAX_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = 148 * 8;
    count_launch();
    if (mode == 0)
        atomic_bench_kernel<0><<<grid, 256, 0, st>>>(region, (unsigned)region_floats, n_ops);
    else if (mode == 1)
        atomic_bench_kernel<1><<<grid, 256, 0, st>>>(region, (unsigned)region_floats, n_ops);
    else if (mode == 2)
        atomic_bench_kernel<2><<<grid, 256, 0, st>>>(region, (unsigned)region_floats, n_ops);
    else if (mode == 3)
        atomic_bench_kernel<3><<<grid, 256, 0, st>>>(region, (unsigned)region_floats, n_ops);
    else if (mode == 4)
        atomic_bench_kernel<4><<<grid, 256, 0, st>>>(region, (unsigned)region_floats, n_ops);
    else
        return CMAX_ERR_BAD_CONFIG;
    return check_launch();
}

int cmax_stage_count(void) { return ST_COUNT; }

const char *cmax_stage_name(int stage)
{
    return (stage >= 0 && stage < ST_COUNT) ? kStageNames[stage] : "";
}

int cmax_stage_timing_enable(int on)
{
    g_timing = on != 0;
    g_ev_used = 0;
    return CMAX_OK;
}

int cmax_stage_timing_read(double *ms_sum, int64_t *count)
{
    if (!ms_sum || !count) return CMAX_ERR_BAD_SHAPE;
    for (int i = 0; i < ST_COUNT; ++i) { ms_sum[i] = 0.0; count[i] = 0; }
    for (int i = 0; i < g_ev_used; ++i) {
        if (cudaEventSynchronize(g_ev[i][1]) != cudaSuccess) return CMAX_ERR_CUDA;
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, g_ev[i][0], g_ev[i][1]) != cudaSuccess) return CMAX_ERR_CUDA;
        ms_sum[g_ev_stage[i]] += ms;
        count[g_ev_stage[i]] += 1;
    }
    g_ev_used = 0;
    return CMAX_OK;
}

int64_t cmax_launch_count(void) { return g_launches; }

int64_t cmax_last_worklist_count(void *stream)
{
    // inspection hook: reads the counter inside the workspace of the most recent forward call, so
    // it is only meaningful while that workspace is still alive
    if (!g_last_work_count) return -1;
    int v = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cudaMemcpyAsync(&v, g_last_work_count, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    return v;
}

int cmax_last_worklist_reasons(int64_t out_host[16], void *stream)
{
    if (!g_last_work_count || !out_host) return CMAX_ERR_WORKSPACE;
    int v[16] = {};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cudaMemcpyAsync(v, g_last_work_count, sizeof(v), cudaMemcpyDeviceToHost, st) != cudaSuccess) return CMAX_ERR_CUDA;
    if (cudaStreamSynchronize(st) != cudaSuccess) return CMAX_ERR_CUDA;
    for (int i = 0; i < 16; ++i) out_host[i] = v[i];
    return CMAX_OK;
}

int cmax_read_status(const void *workspace, int64_t out_host[4], void *stream)
{
    if (!workspace || !out_host) return CMAX_ERR_WORKSPACE;
    long long tmp[4];
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cudaMemcpyAsync(tmp, workspace, sizeof(tmp), cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return CMAX_ERR_CUDA;
    if (cudaStreamSynchronize(st) != cudaSuccess) return CMAX_ERR_CUDA;
    for (int i = 0; i < 4; ++i) out_host[i] = tmp[i];
    return CMAX_OK;
}

}  // extern "C"
