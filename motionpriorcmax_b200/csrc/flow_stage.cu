// flow_stage.cu - dense optical-flow read-out: SURVEY.md section 8(f) rank 4.
//
// Mirrors upstream dense_flow_from_traj (src/utils/flow.py:8-16): the per-trajectory flow list is
// placed on the patch lattice (list_to_grid, src/utils/trajectories.py:54-75) and resized to the
// image with torchvision's BICUBIC + antialias=True, i.e. ATen's separable anti-aliased bicubic
// kernel (A = -0.5): for output index i, centre = scale (i + 0.5), taps
// [max(int(centre - support + 0.5), 0), min(int(centre + support + 0.5), in)) with weights
// cubic((j - centre + 0.5) * invscale) normalised to 1 (border taps are dropped and the rest
// renormalised, unlike the clamping non-antialiased kernel).  Inference / logging path
// (scripts/dsec_inference.py:85-93): forward only.
#include "cmax_common.cuh"

namespace cmax {

__global__ void __launch_bounds__(256)
list_to_grid_kernel(const float *__restrict__ list, const long long *__restrict__ pos, int64_t n,
                    int C, int patch, int hq, int wq, float *__restrict__ grid)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = blockIdx.y;
    if (j >= n) return;
    const long long ky = pos[2 * j] / patch, kx = pos[2 * j + 1] / patch;
    if (ky < 0 || ky >= hq || kx < 0 || kx >= wq) return;
    for (int c = 0; c < C; ++c)
        grid[((b * C + c) * hq + ky) * wq + kx] = list[(b * n + j) * C + c];
}

__device__ __forceinline__ float cubic_aa(float x)
{
    const float a = -0.5f;
    x = fabsf(x);
    if (x < 1.0f) return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
    if (x < 2.0f) return a * (((x - 5.0f) * x + 8.0f) * x - 4.0f);
    return 0.0f;
}

struct Taps { int lo, cnt; float w[8]; };     // up to 8 taps (down-scaling by < 2x fits as well)

__device__ __forceinline__ Taps aa_taps(int i, int in_size, int out_size)
{
    Taps t;
    const float scale = (float)in_size / (float)out_size;
    const float support = scale >= 1.0f ? 2.0f * scale : 2.0f;
    const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
    const float centre = scale * ((float)i + 0.5f);
    t.lo = max((int)(centre - support + 0.5f), 0);
    t.cnt = min(min((int)(centre + support + 0.5f), in_size) - t.lo, 8);
    float tot = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        t.w[k] = k < t.cnt ? cubic_aa(((float)(k + t.lo) - centre + 0.5f) * invscale) : 0.0f;
        tot += t.w[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) t.w[k] = tot != 0.0f ? t.w[k] / tot : 0.0f;
    return t;
}

// ATen resizes the last dimension first, then the rows: keep that order of accumulation
__global__ void __launch_bounds__(256)
resize_bicubic_aa_kernel(const float *__restrict__ src, int hq, int wq, int H, int W,
                         float *__restrict__ dst)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    const int64_t plane = blockIdx.z;
    if (x >= W || y >= H) return;
    const Taps tx = aa_taps(x, wq, W), ty = aa_taps(y, hq, H);
    const float *s = src + plane * (int64_t)hq * wq;
    float acc = 0.0f;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        if (a < ty.cnt) {
            float row = 0.0f;
#pragma unroll
            for (int b = 0; b < 8; ++b)
                if (b < tx.cnt) row += tx.w[b] * __ldg(s + (int64_t)(ty.lo + a) * wq + tx.lo + b);
            acc += ty.w[a] * row;
        }
    }
    dst[plane * (int64_t)H * W + (int64_t)y * W + x] = acc;
}

}  // namespace cmax

using namespace cmax;

extern "C" int cmax_dense_flow(const float *traj_flow, const int64_t *pixel_positions, int64_t B,
                               int64_t n, int32_t C, int32_t patch, int32_t H, int32_t W,
                               float *patch_flow_out, float *dense_out, void *stream)
{
    cmax::DeviceGuard dev_guard(dense_out);
    if (B < 1 || n < 0 || C < 1 || patch < 1 || H < patch || W < patch || !patch_flow_out || !dense_out ||
        (n > 0 && (!traj_flow || !pixel_positions)))
        return CMAX_ERR_BAD_SHAPE;
    if (B > 65535 || B * C > 65535) return CMAX_ERR_UNSUPPORTED;
    const int hq = H / patch, wq = W / patch;                 // flow.py:13-14; hq <= H: up-scaling, <= 5 taps
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(patch_flow_out, 0, sizeof(float) * B * C * hq * wq, st);
    if (n > 0) {
        dim3 g1((unsigned)((n + 255) / 256), (unsigned)B);
        list_to_grid_kernel<<<g1, 256, 0, st>>>(traj_flow, reinterpret_cast<const long long *>(pixel_positions),
                                                 n, C, patch, hq, wq, patch_flow_out);
        count_launch();
    }
    dim3 g2((W + 31) / 32, (H + 7) / 8, (unsigned)(B * C));
    resize_bicubic_aa_kernel<<<g2, dim3(32, 8), 0, st>>>(patch_flow_out, hq, wq, H, W, dense_out);
    count_launch();
    return check_launch();
}
