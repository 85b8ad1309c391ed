// event_stage.cu - per-event work of the CMax loss: LUT lookup + warp (upstream
// src/losses/focus.py:182-195), weights (focus.py:197-214), bilinear vote into the raw IWE
// (src/utils/event_image_converter.py:333-391) and its adjoint (gather of dL/dIWE at the four
// corners, reduction into dL/dLUT).
//
// One pass over the event stream each way: 24 B read per event row, nothing written per
// event.  The reference materialises ~25 [B, M]-sized temporaries - here the
// warped coordinate, weight, corner indices and votes live in registers only.
// IWE / dLUT accumulation: red.global.add.f32 (L2-resident targets), or int64 fixed point
// (2^-32) when deterministic.
#include "cmax_common.cuh"

namespace cmax {

template <bool DET>
__global__ void __launch_bounds__(256)
event_forward_kernel(const float *__restrict__ events, const float *__restrict__ times, Geom g,
                     const float *__restrict__ lut, float *__restrict__ raw,
                     long long *__restrict__ raw_i64, long long *__restrict__ status)
{
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = blockIdx.y;
    if (m >= g.M) return;
    const EventRow e = load_event(events + (b * g.M + m) * 6);
    if (e.valid == 0.0f) return;                      // padding rows vote 0 everywhere
    int64_t cell;
    if (!lut_cell(e, g, b, &cell)) {
        atomicAdd(reinterpret_cast<unsigned long long *>(status), 1ull);
        return;
    }
    const int pol = (g.pab && m >= g.npos) ? 1 : 0;
    const int64_t HW = (int64_t)g.H * g.W;
    for (int r = 0; r < g.R; ++r) {
        const float2 f = __ldg(reinterpret_cast<const float2 *>(lut) + cell * g.R + r);
        const float wy = __fadd_rn(f.x, e.y), wx = __fadd_rn(f.y, e.x);     // focus.py:191
        const float w = event_weight(e, wy, wx, __ldg(times + r), g);
        if (w == 0.0f) continue;
        const Corners c = vote_corners(wy, wx, g.H, g.W);
        const float oy = __fsub_rn(1.0f, c.fy), ox = __fsub_rn(1.0f, c.fx);
        float v[4];
        v[0] = __fmul_rn(__fmul_rn(oy, ox), w);        // event_image_converter.py:382-385
        v[1] = __fmul_rn(__fmul_rn(c.fy, ox), w);
        v[2] = __fmul_rn(__fmul_rn(oy, c.fx), w);
        v[3] = __fmul_rn(__fmul_rn(c.fy, c.fx), w);
        const int64_t base = ((b * g.R + r) * g.P + pol) * HW;
        if (DET) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (c.idx[k] >= 0)
                    atomicAdd(reinterpret_cast<unsigned long long *>(raw_i64 + base + c.idx[k]),
                              (unsigned long long)to_fix(v[k]));
        } else {
            // corners (y, x1) and (y, x1+1) are adjacent floats: one vector red when 8 B aligned
            // (plane base and row pitch are even, so alignment depends on x1 only)
#pragma unroll
            for (int row = 0; row < 2; ++row) {
                const int i0 = c.idx[row], i1 = c.idx[row + 2];
                float *p0 = raw + base + i0;
                const int off = (int)((base + i0) & 3);
                if (i0 >= 0 && i1 >= 0 && (off & 1) == 0) {
                    red_add_f32x2(p0, v[row], v[row + 2]);
                } else if (i0 >= 0 && i1 >= 0 && off == 1) {
                    // (x1, x1+1) sit in the middle of an aligned quad: one 16-byte request
                    red_add_f32x4(p0 - 1, 0.0f, v[row], v[row + 2], 0.0f);
                } else {
                    if (i0 >= 0) atomicAdd(p0, v[row]);
                    if (i1 >= 0) atomicAdd(raw + base + i1, v[row + 2]);
                }
            }
        }
    }
}

// int64 fixed point -> float image
__global__ void fix_to_float_kernel(const long long *__restrict__ in, float *__restrict__ out,
                                    int64_t count)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (float)((double)in[i] * (1.0 / kFixScale));
}

// Backward: g' = w * (d vote / d (wy, wx)) . D   with D = *unscaled* dL/dIWE at the corners;
// dLUT[cell, r] += coef * g'   where coef = grad_loss * (-1 / val^2) / N  (loss.py:12,22-25).
template <bool DET>
__global__ void __launch_bounds__(256)
event_backward_kernel(const float *__restrict__ events, const float *__restrict__ times, Geom g,
                      const float *__restrict__ lut, const float *__restrict__ dimg,
                      const Header *__restrict__ hdr, const float *__restrict__ grad_loss,
                      float *__restrict__ dlut, long long *__restrict__ dlut_i64)
{
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = blockIdx.y;
    if (m >= g.M) return;
    const EventRow e = load_event(events + (b * g.M + m) * 6);
    if (e.valid == 0.0f) return;
    int64_t cell;
    if (!lut_cell(e, g, b, &cell)) return;
    const int pol = (g.pab && m >= g.npos) ? 1 : 0;
    const int64_t HW = (int64_t)g.H * g.W;
    float coef = 1.0f;
    if (!DET) {
        const float val = hdr->val;
        const float N = (float)((double)g.B * g.R * g.P * (double)HW);
        coef = __ldg(grad_loss) * (-(1.0f / (val * val))) / N;
    }
    for (int r = 0; r < g.R; ++r) {
        const float2 f = __ldg(reinterpret_cast<const float2 *>(lut) + cell * g.R + r);
        const float wy = __fadd_rn(f.x, e.y), wx = __fadd_rn(f.y, e.x);
        const float w = event_weight(e, wy, wx, __ldg(times + r), g);
        if (w == 0.0f) continue;
        const Corners c = vote_corners(wy, wx, g.H, g.W);
        const float *D = dimg + ((b * g.R + r) * g.P + pol) * HW;
        const float d00 = c.idx[0] >= 0 ? __ldg(D + c.idx[0]) : 0.0f;
        const float d10 = c.idx[1] >= 0 ? __ldg(D + c.idx[1]) : 0.0f;
        const float d01 = c.idx[2] >= 0 ? __ldg(D + c.idx[2]) : 0.0f;
        const float d11 = c.idx[3] >= 0 ? __ldg(D + c.idx[3]) : 0.0f;
        const float oy = 1.0f - c.fy, ox = 1.0f - c.fx;
        const float gy = w * (ox * (d10 - d00) + c.fx * (d11 - d01));
        const float gx = w * (oy * (d01 - d00) + c.fy * (d11 - d10));
        if (DET) {
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(dlut_i64) + (cell * g.R + r) * 2;
            atomicAdd(dst, (unsigned long long)to_fix(gy));
            atomicAdd(dst + 1, (unsigned long long)to_fix(gx));
        } else {
            red_add_f32x2(dlut + (cell * g.R + r) * 2, coef * gy, coef * gx);
        }
    }
}

// deterministic mode: dLUT = coef * fixed-point sum (+ the smoothness gradient already there)
__global__ void dlut_finalize_kernel(const long long *__restrict__ acc, float *__restrict__ dlut,
                                     const Header *__restrict__ hdr,
                                     const float *__restrict__ grad_loss, double N, int64_t count)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float val = hdr->val;
    const float coef = __ldg(grad_loss) * (-(1.0f / (val * val))) / (float)N;
    dlut[i] += coef * (float)((double)acc[i] * (1.0 / kFixScale));
}

// ---------------------------------------------------------------------------------------------
// stand-alone imager kernels (cmax_create_iwe / cmax_count_image)
// ---------------------------------------------------------------------------------------------
template <int MODE>   // 0: float votes, 1: int64 fixed-point votes, 2: unit counts (int64)
__global__ void __launch_bounds__(256)
splat_kernel(const float *__restrict__ events, const float *__restrict__ weight, int64_t M,
             int64_t stride, int H, int W, int ph, int pw, float *__restrict__ out,
             long long *__restrict__ out_i64)
{
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = blockIdx.y;
    if (m >= M) return;
    const float *row = events + (b * M + m) * stride;
    const float wy = row[0], wx = row[1];
    const float w = weight ? weight[b * M + m] : 1.0f;
    const Corners c = (ph | pw) ? vote_corners_padded(wy, wx, H, W, ph, pw) : vote_corners(wy, wx, H, W);
    const int64_t base = b * (int64_t)H * W;
    if (MODE == 2) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (c.idx[k] >= 0)
                atomicAdd(reinterpret_cast<unsigned long long *>(out_i64 + base + c.idx[k]), 1ull);
        return;
    }
    const float oy = __fsub_rn(1.0f, c.fy), ox = __fsub_rn(1.0f, c.fx);
    float v[4];
    v[0] = __fmul_rn(__fmul_rn(oy, ox), w);
    v[1] = __fmul_rn(__fmul_rn(c.fy, ox), w);
    v[2] = __fmul_rn(__fmul_rn(oy, c.fx), w);
    v[3] = __fmul_rn(__fmul_rn(c.fy, c.fx), w);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (c.idx[k] >= 0) {
            if (MODE == 1)
                atomicAdd(reinterpret_cast<unsigned long long *>(out_i64 + base + c.idx[k]),
                          (unsigned long long)to_fix(v[k]));
            else
                atomicAdd(out + base + c.idx[k], v[k]);
        }
    }
}

int launch_splat(int mode, const float *events, const float *weight, int64_t nb, int64_t M,
                 int64_t stride, int H, int W, int ph, int pw, float *out, long long *out_i64,
                 cudaStream_t st)
{
    if (M == 0 || nb == 0) return CMAX_OK;
    dim3 grid((unsigned)((M + 255) / 256), (unsigned)nb);
    count_launch();
    if (mode == 0)
        splat_kernel<0><<<grid, 256, 0, st>>>(events, weight, M, stride, H, W, ph, pw, out, out_i64);
    else if (mode == 1)
        splat_kernel<1><<<grid, 256, 0, st>>>(events, weight, M, stride, H, W, ph, pw, out, out_i64);
    else
        splat_kernel<2><<<grid, 256, 0, st>>>(events, weight, M, stride, H, W, ph, pw, out, out_i64);
    return check_launch();
}

int launch_fix_to_float(const long long *in, float *out, int64_t count, cudaStream_t st)
{
    if (count == 0) return CMAX_OK;
    fix_to_float_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(in, out, count);
    return check_launch();
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
// phase 0: everything; 1: accumulate only (leaves the int64 sums un-converted when deterministic);
// 2: only the deterministic conversion.  Phases 1 / 2 bracket the all-reduce of the event-sharded mode.
int launch_event_forward(const Geom &g, const Layout &L, const float *events, const float *times,
                         char *ws, cudaStream_t st, int phase)
{
    const int64_t count = g.B * g.R * g.P * (int64_t)g.H * g.W;
    float *raw = reinterpret_cast<float *>(ws + L.raw);
    long long *raw_i64 = reinterpret_cast<long long *>(ws + L.raw_i64);
    long long *status = reinterpret_cast<Header *>(ws + L.header)->status;
    const float *lut = reinterpret_cast<const float *>(ws + L.lut);
    if (phase == 2) return g.det ? launch_fix_to_float(raw_i64, raw, count, st) : CMAX_OK;
    StageScope sc(ST_EVENT_FWD, st);
    count_launch((g.M > 0 ? 1 : 0) + (g.det ? 1 : 0));
    if (g.det)
        cudaMemsetAsync(raw_i64, 0, sizeof(long long) * count, st);
    else
        cudaMemsetAsync(raw, 0, sizeof(float) * count, st);
    if (g.M > 0) {
        dim3 grid((unsigned)((g.M + 255) / 256), (unsigned)g.B);
        if (g.det)
            event_forward_kernel<true><<<grid, 256, 0, st>>>(events, times, g, lut, raw, raw_i64, status);
        else
            event_forward_kernel<false><<<grid, 256, 0, st>>>(events, times, g, lut, raw, raw_i64, status);
    }
    if (g.det && phase == 0) return launch_fix_to_float(raw_i64, raw, count, st);
    return check_launch();
}

int launch_event_backward(const Geom &g, const Layout &L, const float *events, const float *times,
                          const float *grad_loss, char *ws, cudaStream_t st, int phase)
{
    const int64_t count = g.S * g.q * g.R * 2;
    const Header *hdr = reinterpret_cast<const Header *>(ws + L.header);
    const float *lut = reinterpret_cast<const float *>(ws + L.lut);
    const float *dimg = reinterpret_cast<const float *>(ws + L.dimg);
    float *dlut = reinterpret_cast<float *>(ws + L.dlut);
    long long *dlut_i64 = reinterpret_cast<long long *>(ws + L.dlut_i64);
    if (phase == 2) return g.det ? launch_dlut_finalize(g, L, grad_loss, ws, st) : CMAX_OK;
    StageScope sc(ST_EVENT_BWD, st);
    count_launch((g.M > 0 ? 1 : 0) + (g.det ? 1 : 0));
    if (g.det) cudaMemsetAsync(dlut_i64, 0, sizeof(long long) * count, st);
    if (g.M > 0) {
        dim3 grid((unsigned)((g.M + 255) / 256), (unsigned)g.B);
        if (g.det)
            event_backward_kernel<true><<<grid, 256, 0, st>>>(events, times, g, lut, dimg, hdr,
                                                               grad_loss, dlut, dlut_i64);
        else
            event_backward_kernel<false><<<grid, 256, 0, st>>>(events, times, g, lut, dimg, hdr,
                                                                grad_loss, dlut, dlut_i64);
    }
    if (g.det && phase == 0) return launch_dlut_finalize(g, L, grad_loss, ws, st);
    return check_launch();
}

int launch_dlut_finalize(const Geom &g, const Layout &L, const float *grad_loss, char *ws,
                         cudaStream_t st)
{
    const int64_t count = g.S * g.q * g.R * 2;
    const double N = (double)g.B * g.R * g.P * (double)g.H * g.W;
    dlut_finalize_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const long long *>(ws + L.dlut_i64), reinterpret_cast<float *>(ws + L.dlut),
        reinterpret_cast<const Header *>(ws + L.header), grad_loss, N, count);
    return check_launch();
}

}  // namespace cmax
