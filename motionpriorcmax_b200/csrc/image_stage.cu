// image_stage.cu - dense stages of the CMax loss:
//   * 3x3 gaussian blur with reflect padding (torchvision gaussian_blur as called from upstream
//     src/utils/event_image_converter.py:170-175), Sobel with zero padding and the focus
//     functional mean(|dx|+|dy|) / mean(dx^2+dy^2) (src/utils/loss.py:4-27,58-87), fused in one
//     shared-memory tile pass: raw IWE read once, blurred IWE written once, per-CTA partial sum;
//   * the transposed chain (sign / 2x -> Sobel^T -> blur^T incl. the reflect adjoint) producing the
//     *unscaled* dL/dIWE_raw from the raw IWE in one pass with a 4-pixel halo;
//   * Charbonnier smoothness on the LUT or on flow_to_next (focus.py:232-246, loss.py:29-56)
//     and its gradient;
//   * the fixed-order final reduction -> {loss, focus_loss, smoothness_loss}.
#include "cmax_common.cuh"

namespace cmax {

// torchvision _get_gaussian_kernel1d(3, sigma) evaluated in float32 on the device-independent
// host: x = [-1, 0, 1], pdf = exp(-0.5 (x/sigma)^2), normalised.
struct Gauss3 { float g0, g1; };   // [g0, g1, g0]
static Gauss3 gauss3(float sigma)
{
    float e = expf(-0.5f * (1.0f / sigma) * (1.0f / sigma));
    float sum = e + 1.0f + e;
    return Gauss3{e / sum, 1.0f / sum};
}

// reflect-101 index for a 1-pixel pad; anything further out is never used (returns -1)
__device__ __forceinline__ int reflect1(int i, int n)
{
    if (i == -1) return 1;
    if (i == n) return n - 2;
    return (i < 0 || i > n) ? -1 : i;
}

constexpr int T = kImgTile;

// ---- forward: blur + sobel + partial focus sum ------------------------------------------------
__global__ void __launch_bounds__(256)
image_forward_kernel(const float *__restrict__ raw, float *__restrict__ blurred_out, int H, int W,
                     Gauss3 gk, int l2, int variance, double *__restrict__ partials, int n_blocks)
{
    __shared__ float s_raw[T + 4][T + 4 + 1];
    __shared__ float s_blur[T + 2][T + 2 + 1];
    __shared__ double s_red[32];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int64_t plane = blockIdx.z;
    const int y0 = blockIdx.y * T, x0 = blockIdx.x * T;
    const float *img = raw + plane * (int64_t)H * W;
    for (int i = tid; i < (T + 4) * (T + 4); i += 256) {
        int ly = i / (T + 4), lx = i - ly * (T + 4);
        int yy = reflect1(y0 + ly - 2, H), xx = reflect1(x0 + lx - 2, W);
        s_raw[ly][lx] = (yy >= 0 && xx >= 0) ? __ldg(img + (int64_t)yy * W + xx) : 0.0f;
    }
    __syncthreads();
    const float k00 = gk.g0 * gk.g0, k01 = gk.g0 * gk.g1, k11 = gk.g1 * gk.g1;
    for (int i = tid; i < (T + 2) * (T + 2); i += 256) {
        int ly = i / (T + 2), lx = i - ly * (T + 2);
        int yy = y0 + ly - 1, xx = x0 + lx - 1;
        float v = 0.0f;                                    // zero padding of the Sobel input
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            // same summation order as a 3x3 correlation, row-major taps
            v = k00 * s_raw[ly][lx];
            v += k01 * s_raw[ly][lx + 1];
            v += k00 * s_raw[ly][lx + 2];
            v += k01 * s_raw[ly + 1][lx];
            v += k11 * s_raw[ly + 1][lx + 1];
            v += k01 * s_raw[ly + 1][lx + 2];
            v += k00 * s_raw[ly + 2][lx];
            v += k01 * s_raw[ly + 2][lx + 1];
            v += k00 * s_raw[ly + 2][lx + 2];
        }
        s_blur[ly][lx] = v;
    }
    __syncthreads();
    double acc = 0.0, acc2 = 0.0;
    float *outp = blurred_out + plane * (int64_t)H * W;
    for (int i = tid; i < T * T; i += 256) {
        int ly = i / T, lx = i - ly * T;
        int yy = y0 + ly, xx = x0 + lx;
        if (yy < H && xx < W) {
            const float a = s_blur[ly][lx], b = s_blur[ly][lx + 1], c = s_blur[ly][lx + 2];
            const float d = s_blur[ly + 1][lx], f = s_blur[ly + 1][lx + 2];
            const float g_ = s_blur[ly + 2][lx], h = s_blur[ly + 2][lx + 1], k = s_blur[ly + 2][lx + 2];
            const float dx = (c - a) + 2.0f * (f - d) + (k - g_);
            const float dy = (g_ - a) + 2.0f * (h - b) + (k - c);
            const float v = s_blur[ly + 1][lx + 1];
            if (variance) { acc += (double)v; acc2 += (double)v * (double)v; }     // loss.py:14-16
            else acc += l2 ? (double)(dx * dx + dy * dy) : (double)(fabsf(dx) + fabsf(dy));
            outp[(int64_t)yy * W + xx] = v;
        }
    }
    acc = block_sum(acc, s_red);
    if (variance) acc2 = block_sum(acc2, s_red);
    if (tid == 0) {
        const int64_t idx = (plane * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        partials[idx] = acc;
        partials[n_blocks + idx] = acc2;
    }
}

// plain blur for the stand-alone imager
__global__ void __launch_bounds__(256)
blur_kernel(const float *__restrict__ raw, float *__restrict__ out, int H, int W, Gauss3 gk)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    const int64_t plane = blockIdx.z;
    if (x >= W || y >= H) return;
    const float *img = raw + plane * (int64_t)H * W;
    const float kk[3] = {gk.g0, gk.g1, gk.g0};
    float v = 0.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
            v += (kk[a] * kk[b]) * __ldg(img + (int64_t)reflect1(y + a - 1, H) * W + reflect1(x + b - 1, W));
    out[plane * (int64_t)H * W + (int64_t)y * W + x] = v;
}

// ---- backward: unscaled dL/dIWE_raw ------------------------------------------------------------
// D' = blur^T( Sobel_x^T u + Sobel_y^T v ),  (u, v) = (sign dx, sign dy)  [l1]  or (2dx, 2dy) [l2]
// FWD = true: the same pass also emits what image_forward_kernel produces (blurred IWE, per-CTA
// partial focus sum) - D' does not depend on the loss value, so when a backward is known to
// follow (training) one kernel serves both directions and the raw IWE is read once.
template <bool FWD>
__global__ void __launch_bounds__(256)
image_backward_kernel(const float *__restrict__ raw, float *__restrict__ dimg, int H, int W,
                      Gauss3 gk, int l2, int variance, const double *__restrict__ plane_stats,
                      float *__restrict__ blurred_out, double *__restrict__ partials, int n_blocks)
{
    __shared__ double s_red[FWD ? 32 : 1];
    double facc = 0.0;
    __shared__ float s_raw[T + 8][T + 8 + 1];
    __shared__ float s_blur[T + 6][T + 6 + 1];
    __shared__ float s_u[T + 4][T + 4 + 1];
    __shared__ float s_v[T + 4][T + 4 + 1];
    __shared__ float s_g[T + 2][T + 2 + 1];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int64_t plane = blockIdx.z;
    const int y0 = blockIdx.y * T, x0 = blockIdx.x * T;
    const float *img = raw + plane * (int64_t)H * W;
    for (int i = tid; i < (T + 8) * (T + 8); i += 256) {
        int ly = i / (T + 8), lx = i - ly * (T + 8);
        int yy = reflect1(y0 + ly - 4, H), xx = reflect1(x0 + lx - 4, W);
        s_raw[ly][lx] = (yy >= 0 && xx >= 0) ? __ldg(img + (int64_t)yy * W + xx) : 0.0f;
    }
    __syncthreads();
    const float k00 = gk.g0 * gk.g0, k01 = gk.g0 * gk.g1, k11 = gk.g1 * gk.g1;
    for (int i = tid; i < (T + 6) * (T + 6); i += 256) {
        int ly = i / (T + 6), lx = i - ly * (T + 6);
        int yy = y0 + ly - 3, xx = x0 + lx - 3;
        float v = 0.0f;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            v = k00 * s_raw[ly][lx];
            v += k01 * s_raw[ly][lx + 1];
            v += k00 * s_raw[ly][lx + 2];
            v += k01 * s_raw[ly + 1][lx];
            v += k11 * s_raw[ly + 1][lx + 1];
            v += k01 * s_raw[ly + 1][lx + 2];
            v += k00 * s_raw[ly + 2][lx];
            v += k01 * s_raw[ly + 2][lx + 1];
            v += k00 * s_raw[ly + 2][lx + 2];
        }
        s_blur[ly][lx] = v;
    }
    __syncthreads();
    if (variance) {
        // d var / d blurred = 2 (I - mean) / (Np - 1); the event stage multiplies by
        // grad * (-1/val^2) / (planes * Np), so G carries the factor Np / (Np - 1)
        const float mean = (float)plane_stats[2 * plane];
        const float c = 2.0f * (float)((double)H * W / ((double)H * W - 1.0));
        for (int i = tid; i < (T + 2) * (T + 2); i += 256) {
            int ly = i / (T + 2), lx = i - ly * (T + 2);
            int yy = y0 + ly - 1, xx = x0 + lx - 1;
            s_g[ly][lx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? c * (s_blur[ly + 2][lx + 2] - mean) : 0.0f;
        }
    }
    for (int i = tid; i < (variance ? 0 : (T + 4) * (T + 4)); i += 256) {
        int ly = i / (T + 4), lx = i - ly * (T + 4);
        int yy = y0 + ly - 2, xx = x0 + lx - 2;
        float u = 0.0f, v = 0.0f;                         // outside the image: no Sobel output
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const float a = s_blur[ly][lx], b = s_blur[ly][lx + 1], c = s_blur[ly][lx + 2];
            const float d = s_blur[ly + 1][lx], f = s_blur[ly + 1][lx + 2];
            const float g_ = s_blur[ly + 2][lx], h = s_blur[ly + 2][lx + 1], k = s_blur[ly + 2][lx + 2];
            const float dx = (c - a) + 2.0f * (f - d) + (k - g_);
            const float dy = (g_ - a) + 2.0f * (h - b) + (k - c);
            if (l2) { u = 2.0f * dx; v = 2.0f * dy; }
            else {
                u = dx > 0.0f ? 1.0f : (dx < 0.0f ? -1.0f : 0.0f);     // torch sign(0) = 0
                v = dy > 0.0f ? 1.0f : (dy < 0.0f ? -1.0f : 0.0f);
            }
            if (FWD && ly >= 2 && ly < T + 2 && lx >= 2 && lx < T + 2) {      // the tile's own pixels
                facc += l2 ? (double)(dx * dx + dy * dy) : (double)(fabsf(dx) + fabsf(dy));
                blurred_out[plane * (int64_t)H * W + (int64_t)yy * W + xx] = s_blur[ly + 1][lx + 1];
            }
        }
        s_u[ly][lx] = u;
        s_v[ly][lx] = v;
    }
    if (FWD) {
        facc = block_sum(facc, s_red);                      // contains the barriers
        if (tid == 0) {
            const int64_t idx = (plane * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
            partials[idx] = facc;
            partials[n_blocks + idx] = 0.0;
        }
    }
    __syncthreads();
    // G = Sobel_x^T u + Sobel_y^T v on tile + halo 1 (positions outside the image are unused)
    for (int i = tid; i < (variance ? 0 : (T + 2) * (T + 2)); i += 256) {
        int ly = i / (T + 2), lx = i - ly * (T + 2);
        // G[i,j] = sum_{a,b} SX[a,b] u[i-a+1, j-b+1] + SY[a,b] v[i-a+1, j-b+1]
        const float (*U)[T + 4 + 1] = s_u;
        const float (*V)[T + 4 + 1] = s_v;
        const int cy = ly + 1, cx = lx + 1;               // centre in s_u coordinates
        float gx = (U[cy + 1][cx + 1] - U[cy + 1][cx - 1]) + 2.0f * (U[cy][cx + 1] - U[cy][cx - 1]) +
                   (U[cy - 1][cx + 1] - U[cy - 1][cx - 1]);
        // SX^T: u at column j-1 contributes with SX[., 2] = +1/+2/+1, column j+1 with -1/-2/-1
        gx = -gx;
        float gy = (V[cy + 1][cx + 1] - V[cy - 1][cx + 1]) + 2.0f * (V[cy + 1][cx] - V[cy - 1][cx]) +
                   (V[cy + 1][cx - 1] - V[cy - 1][cx - 1]);
        gy = -gy;
        int yy = y0 + ly - 1, xx = x0 + lx - 1;
        s_g[ly][lx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? gx + gy : 0.0f;
    }
    __syncthreads();
    // D'[y,x] = sum over padded positions (u,v) that alias (y,x) of Gp[u,v],
    // Gp[u,v] = sum_{a,b} k[a,b] G[u-a, v-b]   (padded coords = image coords + 1)
    const float kk[3] = {gk.g0, gk.g1, gk.g0};
    float *outp = dimg + plane * (int64_t)H * W;
    for (int i = tid; i < T * T; i += 256) {
        int ly = i / T, lx = i - ly * T;
        int yy = y0 + ly, xx = x0 + lx;
        if (yy >= H || xx >= W) continue;
        // direct term (u, v) = (y + 1, x + 1): s_g is zero outside the image, so the nine taps need
        // no bounds tests
        float acc = 0.0f;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                acc += (kk[a] * kk[b]) * s_g[ly + 2 - a][lx + 2 - b];
        // reflect-padding aliases: padded row 0 -> image row 1, padded row H+1 -> image row H-2
        // (same for columns); only those rows / columns take the generic path
        if (yy == 1 || yy == H - 2 || xx == 1 || xx == W - 2) {
            int us[3], vs[3], nu = 1, nv = 1;
            us[0] = yy + 1;
            vs[0] = xx + 1;
            if (yy == 1) us[nu++] = 0;
            if (yy == H - 2) us[nu++] = H + 1;             // H == 3: both alias row 1
            if (xx == 1) vs[nv++] = 0;
            if (xx == W - 2) vs[nv++] = W + 1;
            for (int iu = 0; iu < nu; ++iu) {
                for (int iv = 0; iv < nv; ++iv) {
                    if (iu == 0 && iv == 0) continue;       // the direct term is already in acc
                    const int u = us[iu], v = vs[iv];
                    for (int a = 0; a < 3; ++a) {
                        const int gy_ = u - a;              // image row of G
                        if (gy_ < 0 || gy_ >= H) continue;
                        for (int b = 0; b < 3; ++b) {
                            const int gx_ = v - b;
                            if (gx_ < 0 || gx_ >= W) continue;
                            acc += (kk[a] * kk[b]) * s_g[gy_ - y0 + 1][gx_ - x0 + 1];
                        }
                    }
                }
            }
        }
        outp[(int64_t)yy * W + xx] = acc;
    }
}

// ---- smoothness -------------------------------------------------------------------------------
// field element (slab-like index f, cell (iy,ix), vector r) of a [F, Hq, Wq, Rv, 2] tensor;
// the reference reshapes to [F*Rv, 2, Hq, Wq] images (focus.py:243-245).
constexpr int ST_W = 32, ST_H = 8;

__global__ void __launch_bounds__(256)
smooth_forward_kernel(const float *__restrict__ field, int Hq, int Wq, int Rv,
                      double *__restrict__ partials)
{
    __shared__ float2 s_f[ST_H + 2][ST_W + 2];
    __shared__ double s_red[32];
    const int tid = threadIdx.y * ST_W + threadIdx.x;
    const int64_t img = blockIdx.z;                // f * Rv + r
    const int64_t f = img / Rv;
    const int r = (int)(img - f * Rv);
    const int y0 = blockIdx.y * ST_H, x0 = blockIdx.x * ST_W;
    const float2 *base = reinterpret_cast<const float2 *>(field) + f * (int64_t)Hq * Wq * Rv + r;
    for (int i = tid; i < (ST_H + 2) * (ST_W + 2); i += 256) {
        int ly = i / (ST_W + 2), lx = i - ly * (ST_W + 2);
        int yy = y0 + ly - 1, xx = x0 + lx - 1;
        s_f[ly][lx] = (yy >= 0 && yy < Hq && xx >= 0 && xx < Wq)
                          ? __ldg(base + ((int64_t)yy * Wq + xx) * Rv)
                          : make_float2(0.f, 0.f);
    }
    __syncthreads();
    double acc = 0.0;
    const int ly = threadIdx.y, lx = threadIdx.x;
    if (y0 + ly < Hq && x0 + lx < Wq) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
#define SF(yy_, xx_) (ch ? s_f[yy_][xx_].y : s_f[yy_][xx_].x)
            const float a = SF(ly, lx), b = SF(ly, lx + 1), c = SF(ly, lx + 2);
            const float d = SF(ly + 1, lx), e = SF(ly + 1, lx + 2);
            const float g_ = SF(ly + 2, lx), h = SF(ly + 2, lx + 1), k = SF(ly + 2, lx + 2);
            const float dx = (c - a) + 2.0f * (e - d) + (k - g_);
            const float dy = (g_ - a) + 2.0f * (h - b) + (k - c);
            acc += (double)sqrtf(dx * dx + kCharbEps2) + (double)sqrtf(dy * dy + kCharbEps2);
#undef SF
        }
    }
    acc = block_sum(acc, s_red);
    if (tid == 0) partials[(img * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = acc;
}

// gradient of  w * (mean(cx) + mean(cy)) / 2  w.r.t. the field, scaled by grad_loss:
//   out = scale * ( SX^T (dx / cx) + SY^T (dy / cy) ),  scale = grad_loss * w / (2 * count)
__global__ void __launch_bounds__(256)
smooth_backward_kernel(const float *__restrict__ field, int Hq, int Wq, int Rv,
                       const float *__restrict__ grad_loss, float w_over_2cnt,
                       float *__restrict__ out)
{
    __shared__ float2 s_f[ST_H + 4][ST_W + 4];
    __shared__ float2 s_u[ST_H + 2][ST_W + 2];     // dx / cx per channel
    __shared__ float2 s_v[ST_H + 2][ST_W + 2];     // dy / cy per channel
    const int tid = threadIdx.y * ST_W + threadIdx.x;
    const int64_t img = blockIdx.z;
    const int64_t f = img / Rv;
    const int r = (int)(img - f * Rv);
    const int y0 = blockIdx.y * ST_H, x0 = blockIdx.x * ST_W;
    const float2 *base = reinterpret_cast<const float2 *>(field) + f * (int64_t)Hq * Wq * Rv + r;
    for (int i = tid; i < (ST_H + 4) * (ST_W + 4); i += 256) {
        int ly = i / (ST_W + 4), lx = i - ly * (ST_W + 4);
        int yy = y0 + ly - 2, xx = x0 + lx - 2;
        s_f[ly][lx] = (yy >= 0 && yy < Hq && xx >= 0 && xx < Wq)
                          ? __ldg(base + ((int64_t)yy * Wq + xx) * Rv)
                          : make_float2(0.f, 0.f);
    }
    __syncthreads();
    for (int i = tid; i < (ST_H + 2) * (ST_W + 2); i += 256) {
        int ly = i / (ST_W + 2), lx = i - ly * (ST_W + 2);
        int yy = y0 + ly - 1, xx = x0 + lx - 1;
        float2 u = make_float2(0.f, 0.f), v = make_float2(0.f, 0.f);
        if (yy >= 0 && yy < Hq && xx >= 0 && xx < Wq) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
#define SF(yy_, xx_) (ch ? s_f[yy_][xx_].y : s_f[yy_][xx_].x)
                const float a = SF(ly, lx), b = SF(ly, lx + 1), c = SF(ly, lx + 2);
                const float d = SF(ly + 1, lx), e = SF(ly + 1, lx + 2);
                const float g_ = SF(ly + 2, lx), h = SF(ly + 2, lx + 1), k = SF(ly + 2, lx + 2);
                const float dx = (c - a) + 2.0f * (e - d) + (k - g_);
                const float dy = (g_ - a) + 2.0f * (h - b) + (k - c);
                const float uu = dx / sqrtf(dx * dx + kCharbEps2);
                const float vv = dy / sqrtf(dy * dy + kCharbEps2);
                if (ch) { u.y = uu; v.y = vv; } else { u.x = uu; v.x = vv; }
#undef SF
            }
        }
        s_u[ly][lx] = u;
        s_v[ly][lx] = v;
    }
    __syncthreads();
    const int ly = threadIdx.y, lx = threadIdx.x;
    const int yy = y0 + ly, xx = x0 + lx;
    if (yy < Hq && xx < Wq) {
        const float scale = __ldg(grad_loss) * w_over_2cnt;
        const int cy = ly + 1, cx = lx + 1;
        float2 res;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
#define SU(yy_, xx_) (ch ? s_u[yy_][xx_].y : s_u[yy_][xx_].x)
#define SV(yy_, xx_) (ch ? s_v[yy_][xx_].y : s_v[yy_][xx_].x)
            float gx = (SU(cy + 1, cx + 1) - SU(cy + 1, cx - 1)) + 2.0f * (SU(cy, cx + 1) - SU(cy, cx - 1)) +
                       (SU(cy - 1, cx + 1) - SU(cy - 1, cx - 1));
            float gy = (SV(cy + 1, cx + 1) - SV(cy - 1, cx + 1)) + 2.0f * (SV(cy + 1, cx) - SV(cy - 1, cx)) +
                       (SV(cy + 1, cx - 1) - SV(cy - 1, cx - 1));
            const float val = scale * (-(gx + gy));
            if (ch) res.y = val; else res.x = val;
#undef SU
#undef SV
        }
        reinterpret_cast<float2 *>(out)[(f * (int64_t)Hq * Wq + (int64_t)yy * Wq + xx) * Rv + r] = res;
    }
}

// ---- final fixed-order reduction ----------------------------------------------------------------
__global__ void __launch_bounds__(256)
finalize_losses_kernel(Header *hdr, const double *__restrict__ fpart, int nf,
                       const double *__restrict__ spart, int ns, double n_pix, double n_smooth,
                       float smooth_w, int variance, int planes, double *__restrict__ plane_stats,
                       float *__restrict__ losses_out)
{
    __shared__ double s_red[32];
    double a = 0.0, b = 0.0;
    if (variance) {
        // per plane: mean and unbiased variance over H*W (torch.var, loss.py:14-16)
        const int per = nf / planes;
        const double np_ = n_pix / planes;
        for (int p = threadIdx.x; p < planes; p += 256) {
            double s1 = 0.0, s2 = 0.0;
            for (int i = 0; i < per; ++i) { s1 += fpart[p * per + i]; s2 += fpart[nf + p * per + i]; }
            plane_stats[2 * p] = s1 / np_;
            plane_stats[2 * p + 1] = (s2 - s1 * s1 / np_) / (np_ - 1.0);
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int p = 0; p < planes; ++p) a += plane_stats[2 * p + 1];     // fixed order
    } else {
        for (int i = threadIdx.x; i < nf; i += 256) a += fpart[i];
        a = block_sum(a, s_red);
    }
    for (int i = threadIdx.x; i < ns; i += 256) b += spart[i];
    b = block_sum(b, s_red);
    if (threadIdx.x == 0) {
        const float val = variance ? (float)(a / planes) : (float)(a / n_pix);   // torch.mean
        const float focus = 1.0f / val;                              // loss.py:12
        float smooth = 0.0f;
        if (ns > 0 && n_smooth > 0) smooth = smooth_w * (float)(b / n_smooth / 2.0);
        hdr->focus_sum = a;
        hdr->smooth_sum = b;
        hdr->val = val;
        hdr->focus = focus;
        hdr->smooth = smooth;
        hdr->loss = focus + smooth;
        losses_out[0] = focus + smooth;
        losses_out[1] = focus;
        losses_out[2] = smooth;
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int launch_blur(const float *raw, float *out, int64_t planes, int H, int W, float sigma,
                cudaStream_t st)
{
    dim3 grid((W + 31) / 32, (H + 7) / 8, (unsigned)planes);
    count_launch();
    blur_kernel<<<grid, dim3(32, 8), 0, st>>>(raw, out, H, W, gauss3(sigma));
    return check_launch();
}

int launch_image_forward(const Geom &g, const Layout &L, char *ws, float *iwes_out, cudaStream_t st)
{
    const int64_t planes = g.B * g.R * g.P;
    dim3 grid((g.W + T - 1) / T, (g.H + T - 1) / T, (unsigned)planes);
    StageScope sc(ST_IMAGE_FWD, st);
    count_launch();
    image_forward_kernel<<<grid, dim3(32, 8), 0, st>>>(
        reinterpret_cast<const float *>(ws + L.raw), iwes_out, g.H, g.W, gauss3(1.0f), g.l2focus,
        g.variance, reinterpret_cast<double *>(ws + L.focus_partials), L.n_img_blocks);
    return check_launch();
}

int launch_image_backward(const Geom &g, const Layout &L, char *ws, cudaStream_t st)
{
    const int64_t planes = g.B * g.R * g.P;
    dim3 grid((g.W + T - 1) / T, (g.H + T - 1) / T, (unsigned)planes);
    StageScope sc(ST_IMAGE_BWD, st);
    count_launch();
    image_backward_kernel<false><<<grid, dim3(32, 8), 0, st>>>(
        reinterpret_cast<const float *>(ws + L.raw), reinterpret_cast<float *>(ws + L.dimg), g.H,
        g.W, gauss3(1.0f), g.l2focus, g.variance, reinterpret_cast<const double *>(ws + L.plane_stats),
        nullptr, nullptr, 0);
    return check_launch();
}

// forward and the focus part of the backward in one pass (training hint; gradient-magnitude
// functional only - the variance functional needs the plane means first)
int launch_image_forward_backward(const Geom &g, const Layout &L, char *ws, float *iwes_out,
                                  cudaStream_t st)
{
    const int64_t planes = g.B * g.R * g.P;
    dim3 grid((g.W + T - 1) / T, (g.H + T - 1) / T, (unsigned)planes);
    StageScope sc(ST_IMAGE_FWD, st);
    count_launch();
    image_backward_kernel<true><<<grid, dim3(32, 8), 0, st>>>(
        reinterpret_cast<const float *>(ws + L.raw), reinterpret_cast<float *>(ws + L.dimg), g.H,
        g.W, gauss3(1.0f), g.l2focus, 0, nullptr, iwes_out,
        reinterpret_cast<double *>(ws + L.focus_partials), L.n_img_blocks);
    return check_launch();
}

static void smooth_field(const Geom &g, const Layout &L, char *ws, const float **field, int64_t *F,
                         int *Rv, size_t *grad_off)
{
    if (g.smooth_next) {
        *field = reinterpret_cast<const float *>(ws + L.f2n);
        *F = g.B * (g.nb - 1);
        *Rv = 1;
        *grad_off = L.df2n;
    } else {
        *field = reinterpret_cast<const float *>(ws + L.lut);
        *F = g.S;
        *Rv = g.R;
        *grad_off = L.dlut;
    }
}

int launch_smooth_forward(const Geom &g, const Layout &L, char *ws, cudaStream_t st)
{
    if (!(g.smooth_w != 0.0f)) return CMAX_OK;
    const float *field;
    int64_t F;
    int Rv;
    size_t go;
    smooth_field(g, L, ws, &field, &F, &Rv, &go);
    if (F == 0) return CMAX_OK;
    dim3 grid((g.Wq + ST_W - 1) / ST_W, (g.Hq + ST_H - 1) / ST_H, (unsigned)(F * Rv));
    StageScope sc(ST_SMOOTH_FWD, st);
    count_launch();
    smooth_forward_kernel<<<grid, dim3(ST_W, ST_H), 0, st>>>(
        field, g.Hq, g.Wq, Rv, reinterpret_cast<double *>(ws + L.smooth_partials));
    return check_launch();
}

int launch_smooth_backward(const Geom &g, const Layout &L, const float *grad_loss, char *ws,
                           cudaStream_t st)
{
    // dLUT must start from the smoothness gradient (on_flow_to_tref) or from zero
    const bool on = g.smooth_w != 0.0f;
    StageScope sc(ST_SMOOTH_BWD, st);
    const size_t dlut_bytes = sizeof(float) * g.S * g.q * g.R * 2;
    if (!on || g.smooth_next) cudaMemsetAsync(ws + L.dlut, 0, dlut_bytes, st);
    if (!on) return check_launch();
    const float *field;
    int64_t F;
    int Rv;
    size_t go;
    smooth_field(g, L, ws, &field, &F, &Rv, &go);
    if (F == 0) return check_launch();
    const double cnt = (double)F * Rv * 2.0 * g.Hq * g.Wq;       // elements of dx (== of dy)
    dim3 grid((g.Wq + ST_W - 1) / ST_W, (g.Hq + ST_H - 1) / ST_H, (unsigned)(F * Rv));
    count_launch();
    smooth_backward_kernel<<<grid, dim3(ST_W, ST_H), 0, st>>>(
        field, g.Hq, g.Wq, Rv, grad_loss, (float)((double)g.smooth_w / (2.0 * cnt)),
        reinterpret_cast<float *>(ws + go));
    return check_launch();
}

int launch_finalize_losses(const Geom &g, const Layout &L, char *ws, float *losses_out,
                           cudaStream_t st)
{
    const double n_pix = (double)g.B * g.R * g.P * (double)g.H * g.W;
    int ns = 0;
    double n_smooth = 0.0;
    if (g.smooth_w != 0.0f) {
        const int64_t F = g.smooth_next ? g.B * (g.nb - 1) : g.S;
        const int Rv = g.smooth_next ? 1 : g.R;
        ns = (int)(F * Rv) * ((g.Wq + ST_W - 1) / ST_W) * ((g.Hq + ST_H - 1) / ST_H);
        n_smooth = (double)F * Rv * 2.0 * g.Hq * g.Wq;
    }
    StageScope sc(ST_FINALIZE, st);
    count_launch();
    finalize_losses_kernel<<<1, 256, 0, st>>>(
        reinterpret_cast<Header *>(ws + L.header),
        reinterpret_cast<const double *>(ws + L.focus_partials), L.n_img_blocks,
        reinterpret_cast<const double *>(ws + L.smooth_partials), ns, n_pix, n_smooth, g.smooth_w,
        g.variance, (int)(g.B * g.R * g.P), reinterpret_cast<double *>(ws + L.plane_stats), losses_out);
    return check_launch();
}

}  // namespace cmax
