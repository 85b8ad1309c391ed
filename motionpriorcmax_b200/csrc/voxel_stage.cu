// voxel_stage.cu - event voxel grid (network input) on the GPU: SURVEY.md section 8(f) rank 3.
//
// Mirrors upstream VoxelGrid.convert (src/loader/dsec/utils.py:29-77): trilinear vote of the
// polarity value 2p-1 into [C, H, W] (8 corners per event, the reference makes 8 masked passes
// with Tensor.put_(accumulate=True) on the CPU inside the DataLoader workers), followed by the
// optional normalisation over the non-zero entries ('mean_std': unbiased std as torch.std;
// 'max').  Same splat pattern as the IWE kernel: one streaming pass, red.global.add.f32 into an
// L2-resident grid (15 x 480 x 640 floats = 18 MB).
// Reference quirks kept on purpose: x0 = int(x) truncates toward zero (not floor), so
// coordinates in (-1, 0) vote with weight 1 - |0 - x| at column 0 and a *negative* weight at
// column 1; t_norm = (C-1) * (t - t[0]) / (t[-1] - t[0]) is evaluated in float32 in that order.
#include "cmax_common.cuh"

namespace cmax {

__global__ void __launch_bounds__(256)
voxel_splat_kernel(const float *__restrict__ x, const float *__restrict__ y,
                   const float *__restrict__ t, const float *__restrict__ p, int64_t n, int C,
                   int H, int W, float *__restrict__ grid)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float t_first = __ldg(t), t_last = __ldg(t + n - 1);
    const float xv = __ldg(x + i), yv = __ldg(y + i), tv = __ldg(t + i);
    const float tn = __fdiv_rn(__fmul_rn((float)(C - 1), __fsub_rn(tv, t_first)), __fsub_rn(t_last, t_first));
    const float value = __fsub_rn(__fmul_rn(2.0f, __ldg(p + i)), 1.0f);
    // Tensor.int(): truncation toward zero; keep the conversion in range for wild inputs
    const int x0 = (int)fminf(fmaxf(truncf(xv), -2.0f), (float)W + 1.0f);
    const int y0 = (int)fminf(fmaxf(truncf(yv), -2.0f), (float)H + 1.0f);
    const int t0 = (int)fminf(fmaxf(truncf(tn), -2.0f), (float)C + 1.0f);
    if (!(tn == tn)) return;                              // 0/0 when all timestamps are equal
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
        const int xl = x0 + dx;
        const float wx = __fsub_rn(1.0f, fabsf(__fsub_rn((float)xl, xv)));
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const int yl = y0 + dy;
            const float wy = __fsub_rn(1.0f, fabsf(__fsub_rn((float)yl, yv)));
#pragma unroll
            for (int dt = 0; dt < 2; ++dt) {
                const int tl = t0 + dt;
                if (xl < W && xl >= 0 && yl < H && yl >= 0 && tl >= 0 && tl < C) {
                    const float wt = __fsub_rn(1.0f, fabsf(__fsub_rn((float)tl, tn)));
                    const float w = __fmul_rn(__fmul_rn(__fmul_rn(value, wx), wy), wt);
                    atomicAdd(grid + ((int64_t)tl * H + yl) * W + xl, w);
                }
            }
        }
    }
}

// stats[0] = sum, stats[1] = count of non-zero entries, stats[2] = max |v|, stats[3] = sum (v-mean)^2
__global__ void __launch_bounds__(256)
voxel_stats_kernel(const float *__restrict__ grid, int64_t count, double *__restrict__ stats, int pass)
{
    __shared__ double s_red[32];
    double a = 0.0, b = 0.0, m = 0.0;
    const double mean = pass == 1 ? (stats[1] > 0.0 ? stats[0] / stats[1] : 0.0) : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __ldg(grid + i);
        if (v != 0.0f) {
            if (pass == 0) { a += (double)v; b += 1.0; m = fmax(m, (double)fabsf(v)); }
            else { const double d = (double)v - mean; a += d * d; }
        }
    }
    a = block_sum(a, s_red);
    b = block_sum(b, s_red);
    // block max through the same tree (max is order independent)
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, s_red[w]);
        if (pass == 0) {
            atomicAdd(stats + 0, a);
            atomicAdd(stats + 1, b);
            // max |v| >= 0: the bit pattern of a non-negative double orders like an integer
            atomicMax(reinterpret_cast<unsigned long long *>(stats + 2), (unsigned long long)__double_as_longlong(m));
        } else {
            atomicAdd(stats + 3, a);
        }
    }
}

// norm_type 1: mean_std over the non-zero entries, 2: divide by max |v|
__global__ void __launch_bounds__(256)
voxel_normalize_kernel(float *__restrict__ grid, int64_t count, const double *__restrict__ stats, int norm_type)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float v = grid[i];
    if (norm_type == 1) {
        const double nnz = stats[1];
        if (nnz > 0.0 && v != 0.0f) {
            const float mean = (float)(stats[0] / nnz);
            const float sd = (float)sqrt(stats[3] / (nnz - 1.0));        // torch.std: unbiased; nnz == 1 -> NaN
            grid[i] = sd > 0.0f ? __fdiv_rn(__fsub_rn(v, mean), sd) : __fsub_rn(v, mean);
        }
    } else if (norm_type == 2) {
        const float mx = (float)stats[2];
        if (mx > 0.0f) grid[i] = __fdiv_rn(v, mx);
    }
}

// quantile clipping (utils.py:56-60): |v| > thr -> sign(v) * thr; thr is read from device memory so
// the caller's order statistic needs no host round trip
__global__ void __launch_bounds__(256)
voxel_clip_kernel(float *__restrict__ grid, int64_t count, const float *__restrict__ thr_ptr)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float thr = __ldg(thr_ptr), v = grid[i];
    if (fabsf(v) > thr) grid[i] = __fmul_rn(v > 0.0f ? 1.0f : -1.0f, thr);
}

static int normalize_grid(float *grid, int64_t count, int norm_type, double *stats, cudaStream_t st)
{
    cudaMemsetAsync(stats, 0, sizeof(double) * 4, st);
    const int blocks = 148 * 8;
    voxel_stats_kernel<<<blocks, 256, 0, st>>>(grid, count, stats, 0);
    if (norm_type == 1) voxel_stats_kernel<<<blocks, 256, 0, st>>>(grid, count, stats, 1);
    voxel_normalize_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(grid, count, stats, norm_type);
    count_launch(norm_type == 1 ? 3 : 2);
    return check_launch();
}

}  // namespace cmax

using namespace cmax;

extern "C" int cmax_voxel_normalize(float *grid, int32_t C, int32_t H, int32_t W, int32_t norm_type,
                                    const float *clip_threshold, double *stats_scratch, void *stream)
{
    cmax::DeviceGuard dev_guard(grid);
    if (C < 1 || H < 1 || W < 1 || !grid) return CMAX_ERR_BAD_SHAPE;
    if ((unsigned)norm_type > 2u) return CMAX_ERR_BAD_CONFIG;
    if (norm_type != 0 && !stats_scratch) return CMAX_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t count = (int64_t)C * H * W;
    if (clip_threshold) {
        voxel_clip_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(grid, count, clip_threshold);
        count_launch();
    }
    if (norm_type != 0) return normalize_grid(grid, count, norm_type, stats_scratch, st);
    return check_launch();
}

extern "C" int cmax_voxel_grid(const float *x, const float *y, const float *t, const float *p,
                               int64_t n, int32_t C, int32_t H, int32_t W, int32_t norm_type,
                               float *grid_out, double *stats_scratch, void *stream)
{
    cmax::DeviceGuard dev_guard(grid_out);
    if (C < 1 || H < 1 || W < 1 || n < 0 || !grid_out || (n > 0 && (!x || !y || !t || !p)))
        return CMAX_ERR_BAD_SHAPE;
    if ((unsigned)norm_type > 2u) return CMAX_ERR_BAD_CONFIG;
    if (norm_type != 0 && !stats_scratch) return CMAX_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t count = (int64_t)C * H * W;
    cudaMemsetAsync(grid_out, 0, sizeof(float) * count, st);
    if (n > 0) {
        voxel_splat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, t, p, n, C, H, W, grid_out);
        count_launch();
    }
    if (norm_type != 0) return normalize_grid(grid_out, count, norm_type, stats_scratch, st);
    return check_launch();
}
