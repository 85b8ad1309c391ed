// lut_stage.cu - flow look-up-table stage of the CMax loss (upstream src/losses/focus.py:115-180).
//
// The reference finds, for every LUT cell centre (q = Hq*Wq queries) and every (sample, bin)
// slab, the K nearest of the n trajectory positions at t_mid[bin] by exhaustive KeOps search
// (n*q distance evaluations per slab), gathers their displacement to each reference time and
// averages.  Here:
//   1. bin_points_kernel     one CTA per slab: counting sort of the n points into a uniform
//                            cell list held in shared memory, runs ordered by trajectory index
//                            (so every later traversal order is deterministic);
//   2. knn_select_kernel     one thread per query: exact K-NN by ring search over the cell
//                            list with a shared-memory max-heap keyed on (distance, index);
//                            emits the LUT (and flow_to_next), and instead of the K indices
//                            only the K-th key (tau, jcut) - 8 B per query instead of 2*K B;
//   3. lut_backward_kernel   one thread per trajectory: *gathers* d loss/d LUT from every query
//                            whose K-set contains it (membership = key <= (tau, jcut),
//                            re-evaluated with bit-identical distance arithmetic): no atomics,
//                            deterministic, and no K-index tensor ever touches HBM.
// Exactness: identical to exhaustive search with "lowest index wins" ties (see oracle).
#include "cmax_common.cuh"

namespace cmax {

// ---------------------------------------------------------------------------------------------
// 1. counting sort into the cell list
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int cell_of(float py, float px, const Geom &g)
{
    float cy = fminf(fmaxf(floorf(py * g.inv_cs), 0.0f), (float)(g.Hc - 1));
    float cx = fminf(fmaxf(floorf(px * g.inv_cs), 0.0f), (float)(g.Wc - 1));
    return (int)cy * g.Wc + (int)cx;     // NaN -> fmaxf picks 0
}

__device__ __forceinline__ const float2 *slab_points(const float *traj, const Geom &g, int64_t slab)
{
    int64_t b = slab / g.nb, bin = slab - b * g.nb;
    return reinterpret_cast<const float2 *>(traj) + ((b * (g.R + g.nb) + g.R + bin) * g.n);
}

__global__ void __launch_bounds__(1024)
bin_points_kernel(const float *__restrict__ traj, Geom g, int *__restrict__ cell_start,
                  float4 *__restrict__ sorted)
{
    extern __shared__ int cnt[];                 // [NC] counters, then cursors
    __shared__ int warp_tot[32];
    const int64_t slab = blockIdx.x;
    const float2 *pts = slab_points(traj, g, slab);
    int *cs_out = cell_start + slab * (g.NC + 1);
    float4 *out = sorted + slab * g.n;
    const int tid = threadIdx.x, nt = blockDim.x;

    for (int c = tid; c < g.NC; c += nt) cnt[c] = 0;
    __syncthreads();
    for (int64_t j = tid; j < g.n; j += nt) {
        float2 p = pts[j];
        atomicAdd(&cnt[cell_of(p.x, p.y, g)], 1);
    }
    __syncthreads();
    // exclusive scan: each thread owns `per` consecutive cells
    const int per = (g.NC + nt - 1) / nt;
    const int c0 = tid * per;
    int local = 0;
    for (int i = 0; i < per; ++i)
        if (c0 + i < g.NC) local += cnt[c0 + i];
    int incl = local;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int v = lane < (nt >> 5) ? warp_tot[lane] : 0;
        int iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += u;
        }
        warp_tot[lane] = iv - v;                 // exclusive warp offsets
    }
    __syncthreads();
    int run = warp_tot[wid] + incl - local;
    for (int i = 0; i < per; ++i) {
        if (c0 + i < g.NC) {
            int v = cnt[c0 + i];
            cnt[c0 + i] = run;
            cs_out[c0 + i] = run;
            run += v;
        }
    }
    if (tid == 0) cs_out[g.NC] = (int)g.n;
    __syncthreads();
    // scatter (cursor = cnt); order inside a cell is fixed afterwards
    for (int64_t j = tid; j < g.n; j += nt) {
        float2 p = pts[j];
        int pos = atomicAdd(&cnt[cell_of(p.x, p.y, g)], 1);
        out[pos] = make_float4(p.x, p.y, __int_as_float((int)j), 0.0f);
    }
    __syncthreads();
    // per-cell insertion sort by trajectory index (runs are a handful of points)
    for (int c = tid; c < g.NC; c += nt) {
        int a = cs_out[c], e = cnt[c];
        if (e - a > 1 && e - a <= 2048) {
            for (int i = a + 1; i < e; ++i) {
                float4 v = out[i];
                int key = __float_as_int(v.z), k = i - 1;
                while (k >= a && __float_as_int(out[k].z) > key) {
                    out[k + 1] = out[k];
                    --k;
                }
                out[k + 1] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 2. exact K-NN selection + LUT interpolation
// ---------------------------------------------------------------------------------------------
struct Heap {                       // per-thread max-heap on (d, j), column `tid` of two smem arrays
    float *hd;
    int *hj;
    int K, cnt;
    float rootd;
    int rootj;
    __device__ __forceinline__ float &D(int k) { return hd[k * kKnnBlock]; }
    __device__ __forceinline__ int &J(int k) { return hj[k * kKnnBlock]; }
    __device__ __forceinline__ void sift_down(int i, int size, float d, int j)
    {
        while (true) {
            int c = 2 * i + 1;
            if (c >= size) break;
            float cd = D(c);
            int cj = J(c);
            if (c + 1 < size) {
                float d2 = D(c + 1);
                int j2 = J(c + 1);
                if (lex_less(cd, cj, d2, j2)) { c = c + 1; cd = d2; cj = j2; }
            }
            if (!lex_less(d, j, cd, cj)) break;
            D(i) = cd;
            J(i) = cj;
            i = c;
        }
        D(i) = d;
        J(i) = j;
    }
    __device__ __forceinline__ void consider(float d, int j)
    {
        if (cnt < K) {
            int i = cnt++;
            while (i > 0) {
                int par = (i - 1) >> 1;
                float pd = D(par);
                int pj = J(par);
                if (!lex_less(pd, pj, d, j)) break;
                D(i) = pd;
                J(i) = pj;
                i = par;
            }
            D(i) = d;
            J(i) = j;
            if (cnt == K) { rootd = D(0); rootj = J(0); }
        } else if (lex_less(d, j, rootd, rootj)) {
            sift_down(0, K, d, j);
            rootd = D(0);
            rootj = J(0);
        }
    }
};

__device__ __forceinline__ void scan_cells(Heap &h, const int *__restrict__ cstart,
                                           const float4 *__restrict__ sorted, int row, int c0,
                                           int c1, const Geom &g, float qy, float qx)
{
    if (row < 0 || row >= g.Hc) return;
    c0 = max(c0, 0);
    c1 = min(c1, g.Wc - 1);
    if (c0 > c1) return;
    int a = __ldg(cstart + row * g.Wc + c0), e = __ldg(cstart + row * g.Wc + c1 + 1);
    for (int i = a; i < e; ++i) {
        float4 r = __ldg(sorted + i);
        h.consider(knn_dist(qy, qx, r.x, r.y, g.l1dist), __float_as_int(r.z));
    }
}

// EMIT: 0 = LUT (+tau/jcut/wsum/f2n) for the loss, 1 = sorted neighbour lists (test entry)
template <int EMIT>
__global__ void __launch_bounds__(kKnnBlock)
knn_select_kernel(const float *__restrict__ traj, Geom g, const int *__restrict__ cell_start,
                  const float4 *__restrict__ sorted_all, float *__restrict__ lut,
                  float *__restrict__ f2n, float *__restrict__ tau, int *__restrict__ jcut,
                  float *__restrict__ wsum, unsigned *__restrict__ tau_max,
                  float *__restrict__ lut_copy, int32_t *__restrict__ ind_out,
                  float *__restrict__ dist_out)
{
    extern __shared__ float heap_mem[];
    __shared__ unsigned blk_max;
    const int tid = threadIdx.x;
    const int64_t slab = blockIdx.y;
    const int tiles_x = (g.Wq + kKnnTileW - 1) / kKnnTileW;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int iy = ty * kKnnTileH + tid / kKnnTileW, ix = tx * kKnnTileW + tid % kKnnTileW;
    const bool active = iy < g.Hq && ix < g.Wq;
    if (tid == 0) blk_max = 0u;
    __syncthreads();

    Heap h;
    h.hd = heap_mem + tid;
    h.hj = reinterpret_cast<int *>(heap_mem + (size_t)g.K * kKnnBlock) + tid;
    h.K = g.K;
    h.cnt = 0;
    h.rootd = INFINITY;
    h.rootj = 0x7fffffff;

    if (active) {
        const int *cstart = cell_start + slab * (g.NC + 1);
        const float4 *sorted = sorted_all + slab * g.n;
        const float qy = __fadd_rn((float)(iy * g.s), g.off);      // focus.py:118-123
        const float qx = __fadd_rn((float)(ix * g.s), g.off);
        const int cqy = min((int)floorf(qy * g.inv_cs), g.Hc - 1);
        const int cqx = min((int)floorf(qx * g.inv_cs), g.Wc - 1);
        int r = g.r0;
        for (int row = cqy - r; row <= cqy + r; ++row)
            scan_cells(h, cstart, sorted, row, cqx - r, cqx + r, g, qy, qx);
        while (true) {
            // lower bound on the distance of every point outside the (2r+1)^2 cell window
            float bnd = INFINITY;
            if (cqy - r > 0) bnd = fminf(bnd, qy - (float)(cqy - r) * g.cs);
            if (cqy + r < g.Hc - 1) bnd = fminf(bnd, (float)(cqy + r + 1) * g.cs - qy);
            if (cqx - r > 0) bnd = fminf(bnd, qx - (float)(cqx - r) * g.cs);
            if (cqx + r < g.Wc - 1) bnd = fminf(bnd, (float)(cqx + r + 1) * g.cs - qx);
            if (bnd == INFINITY) break;                    // window covers the whole grid
            bnd = bnd * (1.0f - 1e-5f);
            if (!g.l1dist) bnd = bnd * bnd;
            if (h.cnt == h.K && h.rootd < bnd) break;      // strict: ties could still win on index
            ++r;
            scan_cells(h, cstart, sorted, cqy - r, cqx - r, cqx + r, g, qy, qx);
            scan_cells(h, cstart, sorted, cqy + r, cqx - r, cqx + r, g, qy, qx);
            for (int row = cqy - r + 1; row <= cqy + r - 1; ++row) {
                scan_cells(h, cstart, sorted, row, cqx - r, cqx - r, g, qy, qx);
                scan_cells(h, cstart, sorted, row, cqx + r, cqx + r, g, qy, qx);
            }
        }
    }

    const int64_t c = (int64_t)iy * g.Wq + ix;
    const int64_t sq = slab * g.q + c;
    if (EMIT == 1) {
        if (active) {
            // heap sort: repeatedly move the current maximum to the end
            for (int m = h.K - 1; m >= 0; --m) {
                float d = h.D(0);
                int j = h.J(0);
                float ld = h.D(m);
                int lj = h.J(m);
                if (m > 0) h.sift_down(0, m, ld, lj);
                ind_out[sq * g.K + m] = j;
                if (dist_out) dist_out[sq * g.K + m] = d;
            }
        }
        return;
    }

    if (active) {
        const int64_t b = slab / g.nb, bin = slab - b * g.nb;
        const float2 *tmid = slab_points(traj, g, slab);
        tau[sq] = h.rootd;
        jcut[sq] = h.rootj;
        atomicMax(&blk_max, __float_as_uint(h.rootd));
        float S = 0.0f;
        if (g.iwd) {
            for (int k = 0; k < g.K; ++k) S = __fadd_rn(S, __fdiv_rn(1.0f, __fadd_rn(h.D(k), kIwdEps)));
            wsum[sq] = S;
        }
        const float Kf = (float)g.K;
        for (int r = 0; r < g.R; ++r) {
            const float2 *tref = reinterpret_cast<const float2 *>(traj) + (b * (g.R + g.nb) + r) * g.n;
            float ay = 0.0f, ax = 0.0f;
            for (int k = 0; k < g.K; ++k) {
                int j = h.J(k);
                float2 pr = __ldg(tref + j), pm = __ldg(tmid + j);
                float fy = __fsub_rn(pr.x, pm.x), fx = __fsub_rn(pr.y, pm.y);     // focus.py:141
                if (g.iwd) {
                    float w = __fdiv_rn(__fdiv_rn(1.0f, __fadd_rn(h.D(k), kIwdEps)), S);
                    fy = __fmul_rn(w, fy);
                    fx = __fmul_rn(w, fx);
                }
                ay = __fadd_rn(ay, fy);
                ax = __fadd_rn(ax, fx);
            }
            if (!g.iwd) { ay = __fdiv_rn(ay, Kf); ax = __fdiv_rn(ax, Kf); }     // torch.mean
            float2 v = make_float2(ay, ax);
            reinterpret_cast<float2 *>(lut)[sq * g.R + r] = v;
            if (lut_copy) reinterpret_cast<float2 *>(lut_copy)[sq * g.R + r] = v;
        }
        if (f2n != nullptr && bin < g.nb - 1) {                     // focus.py:170-176
            const float2 *tnext = tmid + g.n;
            float ay = 0.0f, ax = 0.0f;
            for (int k = 0; k < g.K; ++k) {
                int j = h.J(k);
                float2 pn = __ldg(tnext + j), pm = __ldg(tmid + j);
                ay = __fadd_rn(ay, __fsub_rn(pn.x, pm.x));
                ax = __fadd_rn(ax, __fsub_rn(pn.y, pm.y));
            }
            reinterpret_cast<float2 *>(f2n)[(b * (g.nb - 1) + bin) * g.q + c] =
                make_float2(__fdiv_rn(ay, Kf), __fdiv_rn(ax, Kf));
        }
    }
    __syncthreads();
    if (tid == 0 && blk_max != 0u) atomicMax(tau_max + slab, blk_max);
}

// ---------------------------------------------------------------------------------------------
// 3. backward: gather d loss / d LUT into the trajectories
// ---------------------------------------------------------------------------------------------
// dtraj[b, r, j]      =  sum_bins sum_{c : j in KNN(b,bin,c)} w(c,j) dLUT[b,bin,c,r]
// dtraj[b, R+bin, j]  = -sum_r (same inner sum) [- / + the flow_to_next terms]
template <int RT>   // RT = compile-time R (1) or 0 for the generic loop
__global__ void __launch_bounds__(128)
lut_backward_kernel(const float *__restrict__ traj, Geom g, const float *__restrict__ tau,
                    const int *__restrict__ jcut, const float *__restrict__ wsum,
                    const unsigned *__restrict__ tau_max, const float *__restrict__ dlut,
                    const float *__restrict__ df2n, float *__restrict__ dtraj)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = blockIdx.y;
    if (j >= g.n) return;
    const int R = RT ? RT : g.R;
    float2 accr[RT ? RT : kMaxTref];
#pragma unroll
    for (int r = 0; r < (RT ? RT : kMaxTref); ++r) accr[r] = make_float2(0.f, 0.f);
    float2 carry = make_float2(0.f, 0.f);
    const float invK = 1.0f / (float)g.K;
    const float fs = (float)g.s;
    float2 *dt = reinterpret_cast<float2 *>(dtraj) + b * (g.R + g.nb) * g.n;

    for (int bin = 0; bin < g.nb; ++bin) {
        const int64_t slab = b * g.nb + bin;
        const float2 p = slab_points(traj, g, slab)[j];
        float tm = __uint_as_float(tau_max[slab]);
        float rho = g.l1dist ? tm : sqrtf(tm);
        rho = rho * 1.0001f + 1e-3f;
        float2 mid = make_float2(0.f, 0.f);      // sum_r of this bin's gather
        float2 nxt = make_float2(0.f, 0.f);      // flow_to_next gather of this bin
        float2 binr[RT ? RT : kMaxTref];
#pragma unroll
        for (int r = 0; r < (RT ? RT : kMaxTref); ++r) binr[r] = make_float2(0.f, 0.f);
        if (p.x == p.x && p.y == p.y && rho == rho) {
            int iy0 = max(0, (int)floorf((p.x - rho - g.off) / fs));
            int iy1 = min(g.Hq - 1, (int)ceilf((p.x + rho - g.off) / fs));
            int ix0 = max(0, (int)floorf((p.y - rho - g.off) / fs));
            int ix1 = min(g.Wq - 1, (int)ceilf((p.y + rho - g.off) / fs));
            for (int iy = iy0; iy <= iy1; ++iy) {
                const float qy = __fadd_rn((float)(iy * g.s), g.off);
                for (int ix = ix0; ix <= ix1; ++ix) {
                    const float qx = __fadd_rn((float)(ix * g.s), g.off);
                    const float d = knn_dist(qy, qx, p.x, p.y, g.l1dist);
                    const int64_t sq = slab * g.q + (int64_t)iy * g.Wq + ix;
                    const float tc = __ldg(tau + sq);
                    if (d < tc || (d == tc && (int)j <= __ldg(jcut + sq))) {
                        float w = invK;
                        if (g.iwd)
                            w = __fdiv_rn(__fdiv_rn(1.0f, __fadd_rn(d, kIwdEps)), __ldg(wsum + sq));
                        const float2 *dl = reinterpret_cast<const float2 *>(dlut) + sq * g.R;
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            float2 v = __ldg(dl + r);
                            binr[r].x += w * v.x;
                            binr[r].y += w * v.y;
                        }
                        if (df2n != nullptr && bin < g.nb - 1) {
                            float2 v = __ldg(reinterpret_cast<const float2 *>(df2n) +
                                             (b * (g.nb - 1) + bin) * g.q + (int64_t)iy * g.Wq + ix);
                            nxt.x += invK * v.x;
                            nxt.y += invK * v.y;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            accr[r].x += binr[r].x;
            accr[r].y += binr[r].y;
            mid.x += binr[r].x;
            mid.y += binr[r].y;
        }
        dt[(g.R + bin) * g.n + j] = make_float2(-mid.x - nxt.x + carry.x, -mid.y - nxt.y + carry.y);
        carry = nxt;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) dt[r * g.n + j] = accr[r];
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int launch_lut_forward(const Geom &g, const Layout &L, const float *traj, char *ws,
                       float *flow_lut_out, int32_t *ind_out, float *dist_out, cudaStream_t st)
{
    int *cell_start = reinterpret_cast<int *>(ws + L.cell_start);
    float4 *sorted = reinterpret_cast<float4 *>(ws + L.sorted);
    size_t smem_bin = (size_t)g.NC * sizeof(int);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(bin_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kMaxCells * (int)sizeof(int));
        cudaFuncSetAttribute(knn_select_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kMaxKnn * kKnnBlock * 8);
        cudaFuncSetAttribute(knn_select_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kMaxKnn * kKnnBlock * 8);
        attr_done = true;
    }
    {
        StageScope sc(ST_BIN_POINTS, st);
        bin_points_kernel<<<(unsigned)g.S, 1024, smem_bin, st>>>(traj, g, cell_start, sorted);
        count_launch();
    }
    const int tiles = ((g.Wq + kKnnTileW - 1) / kKnnTileW) * ((g.Hq + kKnnTileH - 1) / kKnnTileH);
    dim3 grid(tiles, (unsigned)g.S);
    size_t smem_heap = (size_t)g.K * kKnnBlock * 8;
    StageScope sc(ST_KNN_SELECT, st);
    count_launch();
    if (ind_out != nullptr) {
        knn_select_kernel<1><<<grid, kKnnBlock, smem_heap, st>>>(
            traj, g, cell_start, sorted, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
            nullptr, ind_out, dist_out);
        return check_launch();
    }
    unsigned *tau_max = reinterpret_cast<unsigned *>(ws + L.tau_max);
    cudaMemsetAsync(tau_max, 0, sizeof(unsigned) * g.S, st);
    const bool want_next = g.smooth_next && g.smooth_w > 0.0f && g.nb > 1;
    knn_select_kernel<0><<<grid, kKnnBlock, smem_heap, st>>>(
        traj, g, cell_start, sorted, reinterpret_cast<float *>(ws + L.lut),
        want_next ? reinterpret_cast<float *>(ws + L.f2n) : nullptr,
        reinterpret_cast<float *>(ws + L.tau), reinterpret_cast<int *>(ws + L.jcut),
        reinterpret_cast<float *>(ws + L.wsum), tau_max, flow_lut_out, nullptr, nullptr);
    return check_launch();
}

int launch_lut_backward(const Geom &g, const Layout &L, const float *traj, char *ws,
                        float *dtraj, cudaStream_t st)
{
    dim3 grid((unsigned)((g.n + 127) / 128), (unsigned)g.B);
    const bool want_next = g.smooth_next && g.smooth_w > 0.0f && g.nb > 1;
    const float *tau = reinterpret_cast<const float *>(ws + L.tau);
    const int *jcut = reinterpret_cast<const int *>(ws + L.jcut);
    const float *wsum = reinterpret_cast<const float *>(ws + L.wsum);
    const unsigned *tmax = reinterpret_cast<const unsigned *>(ws + L.tau_max);
    StageScope sc(ST_LUT_BWD, st);
    count_launch();
    const float *dlut = reinterpret_cast<const float *>(ws + L.dlut);
    const float *df2n = want_next ? reinterpret_cast<const float *>(ws + L.df2n) : nullptr;
    if (g.R == 1)
        lut_backward_kernel<1><<<grid, 128, 0, st>>>(traj, g, tau, jcut, wsum, tmax, dlut, df2n, dtraj);
    else
        lut_backward_kernel<0><<<grid, 128, 0, st>>>(traj, g, tau, jcut, wsum, tmax, dlut, df2n, dtraj);
    return check_launch();
}

}  // namespace cmax
