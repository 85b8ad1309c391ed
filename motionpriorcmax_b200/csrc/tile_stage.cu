// tile_stage.cu - the event stage on the PACKED, tile-binned event layout (the
// north star's "bin events by spatial tile, accumulate privately in shared memory, flush once").
//
// Same arithmetic as event_stage.cu (upstream src/losses/focus.py:182-230 and
// src/utils/event_image_converter.py:333-391), different data movement:
//
//   packed records  float4 [B, M]   (y, x, t, meta)  meta = bin << 24 | iy << 12 | ix, the event's
//                                    LUT cell (focus.py:185-187) evaluated ONCE at pack time;
//                                    only valid rows are kept (16 B instead of 24 B, no padding)
//   seg_start       int32  [B, G*NT + 1]  records of a sample are grouped by polarity group g
//                                    (G = 2 when polarity aware) and by SOURCE tile T of ct x ct
//                                    LUT cells (~32 x 32 pixels); prefix offsets per sample.
//
// Events of one (tile, group) segment are warped by the flows of one small LUT block, so their
// votes land in a compact neighbourhood of the tile.  One CTA per segment (slice):
//   forward   votes go into a 64 x 64 pixel window in SHARED memory whose origin follows the flow
//             range of the tile's LUT block; one flush of the non-zero quads with
//             red.global.add.v4.f32 (u64 atomics when deterministic).  The window accumulates in
//             2^-32 FIXED POINT as (low, high) 32-bit words with native ATOMS.ADD and an explicit
//             carry: float and 64-bit shared-memory atomics compile to compare-and-swap spin
//             loops on sm_100 (measured 2x slower), integer 32-bit adds are native and exact.
//             Votes that leave the window fall back to the global reds of event_stage.cu, so any
//             flow magnitude stays exact.
//   backward  the same window of dL/dIWE is staged in shared memory, the four gathers are
//             shared-memory loads, and (g_y, g_x) is reduced per LUT cell of the tile in shared
//             memory before ONE red.global.add.v2.f32 per touched cell.
// The binning itself does not depend on the flow: it is done once per window by the loader
// (motionpriorcmax_b200.io.pack_events_host) or on the device (pack_* kernels below).
#include <type_traits>

#include "cmax_common.cuh"

namespace cmax {

constexpr int kWin = 64;                 // shared-memory window edge in pixels
constexpr int kTileThreads = 256;
constexpr int kMaxSmemAcc = 96 * 1024;   // dLUT block accumulated in shared memory up to this size

__device__ __forceinline__ float4 ld_stream_f4(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// 64-bit fixed-point add into a (low, high) pair of shared-memory words: two native 32-bit
// atomics; the carry out of the low word is recovered from the value the first add returns.
__device__ __forceinline__ void smem_add_fix(unsigned *lo, unsigned *hi, long long x)
{
    if (x == 0) return;
    const unsigned xl = (unsigned)x, xh = (unsigned)((unsigned long long)x >> 32);
    const unsigned old = atomicAdd(lo, xl);
    const unsigned h = xh + ((old + xl) < old ? 1u : 0u);
    if (h) atomicAdd(hi, h);
}

// Float mode: ONE 32-bit word per window pixel, votes in 2^-24 fixed point (a vote of the packed
// layout is a bilinear weight in [0, 1]; its float32 rounding error is of the same size).  A word
// wraps at 256.0: the wrap is seen in the value the atomic returns and carried to the global image
// as one red of 256.0 - rare (dense edges), exact.  Half the ATOMS, half the shared memory, half
// the zeroing and flushing of the two-word int64 form the deterministic mode keeps.
constexpr float kFloatFixScale = 16777216.0f;        // 2^24
__device__ __forceinline__ void smem_add_vote(unsigned *word, float v, float *carry_to)
{
    if (v > 0.0f) {
        const unsigned x = __float2uint_rn(__fmul_rn(v, kFloatFixScale));
        if (x) {
            const unsigned old = atomicAdd(word, x);
            if (old + x < old) atomicAdd(carry_to, 256.0f);
        }
    } else if (v < 0.0f) {
        atomicAdd(carry_to, v);      // a fractional part of -1e-6 (event_image_converter.py:357): straight to HBM
    }
}

// TMA reduce-add of a contiguous shared-memory run into global memory (float32): the flush of one
// window row is one asynchronous bulk operation instead of 16 REDG.v4 issued by as many threads.
__device__ __forceinline__ void bulk_reduce_add_f32(float *gdst, const float *ssrc, unsigned bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst),
                 "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ long long smem_get_fix(const unsigned *lo, const unsigned *hi)
{
    return (long long)(((unsigned long long)*hi << 32) | (unsigned long long)*lo);
}

// ---------------------------------------------------------------------------------------------
// device-side packing: count -> scan -> scatter
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pack_key(const EventRow &e, const Geom &g, int64_t m, int *key,
                                         unsigned *meta)
{
    int64_t cell;
    if (!lut_cell(e, g, 0, &cell)) return false;
    const int bin = (int)(cell / g.q);
    const int rem = (int)(cell - (int64_t)bin * g.q);
    const int iy = rem / g.Wq, ix = rem - iy * g.Wq;
    const int grp = (g.pab && m >= g.npos) ? 1 : 0;
    *key = grp * g.nt + (iy / g.ct) * g.ntx + ix / g.ct;
    *meta = ((unsigned)bin << 24) | ((unsigned)iy << 12) | (unsigned)ix;
    return true;
}

constexpr int kPackThreads = 512;
constexpr int kPackRows = 8;                            // rows per thread: a CTA bins 4096 rows

// Histogram of one 4096-row chunk in shared memory, then one global add per non-empty segment
// (a global atomic per event serialises on the few hundred counters of a sample).
__global__ void __launch_bounds__(kPackThreads)
pack_count_kernel(const float *__restrict__ events, Geom g, int *__restrict__ counts,
                  long long *__restrict__ skipped)
{
    extern __shared__ int s_hist[];
    const int nkeys = g.P * g.nt;
    const int64_t b = blockIdx.y;
    const int tid = threadIdx.x;
    for (int k = tid; k < nkeys; k += kPackThreads) s_hist[k] = 0;
    __syncthreads();
    int n_out = 0, n_odd = 0;
#pragma unroll
    for (int u = 0; u < kPackRows; ++u) {
        const int64_t m = ((int64_t)blockIdx.x * kPackRows + u) * kPackThreads + tid;
        if (m >= g.M) continue;
        const EventRow e = load_event(events + (b * g.M + m) * 6);
        if (e.valid == 0.0f) continue;
        int key;
        unsigned meta;
        if (!pack_key(e, g, m, &key, &meta)) { ++n_out; continue; }
        if (e.valid != 1.0f) ++n_odd;
        atomicAdd(&s_hist[key], 1);
    }
    __syncthreads();
    for (int k = tid; k < nkeys; k += kPackThreads) {
        const int c = s_hist[k];
        if (c) atomicAdd(counts + b * (int64_t)nkeys + k, c);
    }
    if (skipped) {
        if (n_out) atomicAdd(reinterpret_cast<unsigned long long *>(skipped), (unsigned long long)n_out);
        if (n_odd) atomicAdd(reinterpret_cast<unsigned long long *>(skipped + 1), (unsigned long long)n_odd);
    }
}

// one CTA per sample: exclusive scan of the G*NT counters -> seg_start, cursors
__global__ void __launch_bounds__(1024)
pack_scan_kernel(Geom g, int *__restrict__ counts, int *__restrict__ seg_start)
{
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int n = g.P * g.nt;
    const int64_t b = blockIdx.x;
    int *cnt = counts + b * (int64_t)n;
    int *out = seg_start + b * (int64_t)(n + 1);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const int v = i < n ? cnt[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = warp_tot[lane], iw = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int u = __shfl_up_sync(0xffffffffu, iw, o);
                if (lane >= o) iw += u;
            }
            warp_tot[lane] = iw - w;
        }
        __syncthreads();
        const int excl = carry + warp_tot[wid] + incl - v;
        if (i < n) {
            out[i] = excl;
            cnt[i] = excl;                               // cursor for the scatter
        }
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry;
}

// Same chunking: local ranks from the shared-memory histogram, one global cursor reservation per
// non-empty segment of the chunk, then the records are written to their final slots.
__global__ void __launch_bounds__(kPackThreads)
pack_scatter_kernel(const float *__restrict__ events, Geom g, int *__restrict__ cursor,
                    float4 *__restrict__ records)
{
    extern __shared__ int s_hist[];
    const int nkeys = g.P * g.nt;
    const int64_t b = blockIdx.y;
    const int tid = threadIdx.x;
    for (int k = tid; k < nkeys; k += kPackThreads) s_hist[k] = 0;
    __syncthreads();
    float ry[kPackRows], rx[kPackRows], rt[kPackRows];
    int key[kPackRows], rank[kPackRows];
    unsigned meta[kPackRows];
#pragma unroll
    for (int u = 0; u < kPackRows; ++u) {
        key[u] = -1;
        const int64_t m = ((int64_t)blockIdx.x * kPackRows + u) * kPackThreads + tid;
        if (m >= g.M) continue;
        const EventRow e = load_event(events + (b * g.M + m) * 6);
        if (e.valid == 0.0f) continue;
        int k;
        if (!pack_key(e, g, m, &k, &meta[u])) continue;
        key[u] = k;
        ry[u] = e.y; rx[u] = e.x; rt[u] = e.t;
        rank[u] = atomicAdd(&s_hist[k], 1);
    }
    __syncthreads();
    for (int k = tid; k < nkeys; k += kPackThreads) {
        const int c = s_hist[k];
        if (c) s_hist[k] = atomicAdd(cursor + b * (int64_t)nkeys + k, c);
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kPackRows; ++u)
        if (key[u] >= 0)
            records[b * g.M + s_hist[key[u]] + rank[u]] =
                make_float4(ry[u], rx[u], rt[u], __uint_as_float(meta[u]));
}

// Run lookup for the expand kernels.  A per-thread binary search of the window's run table is a
// chain of 14 dependent loads (measured: 0.29 ms for a 15 M-event batch, three times the memory
// time).  Instead run_index_kernel notes, for every block of 256 records, the run its first record
// sits in (one thread per run: a run covers at most a few block boundaries), and the threads of the
// expand kernel walk forward from there - a handful of L1-resident loads.
constexpr int kExpandBlock = 256;

__global__ void __launch_bounds__(256)
run_index_kernel(const int *__restrict__ fine_start, int F, int64_t blocks_per_window, int *__restrict__ block_run)
{
    const int b = blockIdx.y;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int *fs = fine_start + (int64_t)b * (F + 1);
    const int a = __ldg(fs + f), e = __ldg(fs + f + 1);
    if (e <= a) return;
    for (int64_t c = (a + kExpandBlock - 1) / kExpandBlock; c * kExpandBlock < e && c < blocks_per_window; ++c)
        block_run[b * blocks_per_window + c] = f;          // record c * 256 lies in run f
}

__device__ __forceinline__ int run_of(const int *__restrict__ fs, int F, int i, int start_run)
{
    int lo = start_run;
    while (lo + 1 < F && __ldg(fs + lo + 1) <= i) ++lo;     // passes over empty runs too
    return lo;
}

// ---------------------------------------------------------------------------------------------
// compact wire layout (12 B per event, host packed) -> packed records + seg_start
// ---------------------------------------------------------------------------------------------
// One thread per record: the run it sits in (binary search in the window's fine_start table, which
// stays in L1 / L2: 36 KB per DSEC window) gives the time bin, (y, x) give the LUT cell with the
// reference's own float floor division (focus.py:186-187), exactly what the packers store in `meta`.
__global__ void __launch_bounds__(256)
expand_compact_kernel(const float *__restrict__ coords, const int *__restrict__ fine_start,
                      const long long *__restrict__ sample_off, const int *__restrict__ block_run,
                      int64_t blocks_per_window, Geom g, int64_t Mp, float4 *__restrict__ records,
                      int *__restrict__ seg_start)
{
    const int b = blockIdx.y;
    const int F = g.P * g.nt * g.nb;
    const int *fs = fine_start + (int64_t)b * (F + 1);
    const int count = __ldg(fs + F);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= g.P * g.nt)                                   // coarse table: every nb-th fine entry
        seg_start[(int64_t)b * (g.P * g.nt + 1) + i] = __ldg(fs + i * g.nb);
    if (i >= count) return;
    const int lo = run_of(fs, F, (int)i, __ldg(block_run + b * blocks_per_window + blockIdx.x));
    const int bin = lo % g.nb;
    const float *c = coords + (__ldg(sample_off + b) + i) * 3;
    const float y = __ldg(c), x = __ldg(c + 1), t = __ldg(c + 2);
    const float fsz = (float)g.s;
    const int iy = (int)floordiv_f32(y, fsz), ix = (int)floordiv_f32(x, fsz);
    const unsigned meta = ((unsigned)bin << 24) | ((unsigned)iy << 12) | (unsigned)ix;
    records[(int64_t)b * Mp + i] = make_float4(y, x, t, __uint_as_float(meta));
}

int launch_expand_compact(const Geom &g, const float *coords, const int *fine_start,
                          const long long *sample_off, int64_t Mp, float4 *records, int *seg_start,
                          int *block_run, cudaStream_t st)
{
    const int64_t span = Mp > g.P * g.nt + 1 ? Mp : g.P * g.nt + 1;
    const int64_t nblk = (span + kExpandBlock - 1) / kExpandBlock;
    const int F = g.P * g.nt * g.nb;
    dim3 grid((unsigned)nblk, (unsigned)g.B);
    StageScope sc(ST_PACK, st);
    count_launch(2);
    run_index_kernel<<<dim3((unsigned)((F + 255) / 256), (unsigned)g.B), 256, 0, st>>>(fine_start, F, nblk, block_run);
    expand_compact_kernel<<<grid, kExpandBlock, 0, st>>>(coords, fine_start, sample_off, block_run, nblk, g, Mp,
                                                         records, seg_start);
    return check_launch();
}

// ---------------------------------------------------------------------------------------------
// bit-packed wire layout (host_pack.cpp: fixed-width bit-pattern deltas per run) -> packed records
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned bit_field(const unsigned *__restrict__ words, unsigned long long bit, int w)
{
    if (w == 0) return 0u;
    const unsigned long long wi = bit >> 5;
    const unsigned v = __funnelshift_r(__ldg(words + wi), __ldg(words + wi + 1), (unsigned)(bit & 31u));
    return w >= 32 ? v : (v & ((1u << w) - 1u));
}

__global__ void __launch_bounds__(256)
expand_bitpacked_kernel(const unsigned *__restrict__ words_all, const int *__restrict__ fine_start,
                        const uint4 *__restrict__ run_hdr, const int *__restrict__ run_word,
                        const long long *__restrict__ word_off, const int *__restrict__ block_run,
                        int64_t blocks_per_window, Geom g, int64_t Mp, float4 *__restrict__ records,
                        int *__restrict__ seg_start)
{
    const int b = blockIdx.y;
    const int F = g.P * g.nt * g.nb;
    const int *fs = fine_start + (int64_t)b * (F + 1);
    const int count = __ldg(fs + F);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= g.P * g.nt) seg_start[(int64_t)b * (g.P * g.nt + 1) + i] = __ldg(fs + i * g.nb);
    if (i >= count) return;
    const int lo = run_of(fs, F, (int)i, __ldg(block_run + b * blocks_per_window + blockIdx.x));
    const uint4 h = __ldg(run_hdr + (int64_t)b * F + lo);
    const int wy = (int)(h.w & 255u), wx = (int)((h.w >> 8) & 255u), wt = (int)((h.w >> 16) & 255u);
    const unsigned *words = words_all + __ldg(word_off + b);
    unsigned long long bit = (unsigned long long)__ldg(run_word + (int64_t)b * (F + 1) + lo) * 32ull +
                             (unsigned long long)((int)i - __ldg(fs + lo)) * (unsigned)(wy + wx + wt);
    const float y = __uint_as_float(h.x + bit_field(words, bit, wy));
    bit += (unsigned)wy;
    const float x = __uint_as_float(h.y + bit_field(words, bit, wx));
    bit += (unsigned)wx;
    const float t = __uint_as_float(h.z + bit_field(words, bit, wt));
    const int bin = lo % g.nb;
    const float fsz = (float)g.s;
    const int iy = (int)floordiv_f32(y, fsz), ix = (int)floordiv_f32(x, fsz);
    const unsigned meta = ((unsigned)bin << 24) | ((unsigned)iy << 12) | (unsigned)ix;
    records[(int64_t)b * Mp + i] = make_float4(y, x, t, __uint_as_float(meta));
}

int launch_expand_bitpacked(const Geom &g, const unsigned *words, const int *fine_start, const unsigned *run_hdr,
                            const int *run_word, const long long *word_off, int64_t Mp, float4 *records,
                            int *seg_start, int *block_run, cudaStream_t st)
{
    const int64_t span = Mp > g.P * g.nt + 1 ? Mp : g.P * g.nt + 1;
    const int64_t nblk = (span + kExpandBlock - 1) / kExpandBlock;
    const int F = g.P * g.nt * g.nb;
    dim3 grid((unsigned)nblk, (unsigned)g.B);
    StageScope sc(ST_PACK, st);
    count_launch(2);
    run_index_kernel<<<dim3((unsigned)((F + 255) / 256), (unsigned)g.B), 256, 0, st>>>(fine_start, F, nblk, block_run);
    expand_bitpacked_kernel<<<grid, kExpandBlock, 0, st>>>(words, fine_start, reinterpret_cast<const uint4 *>(run_hdr),
                                                           run_word, word_off, block_run, nblk, g, Mp, records,
                                                           seg_start);
    return check_launch();
}

int launch_pack_events(const Geom &g, const float *events, float4 *records, int *seg_start,
                       int *scratch, long long *skipped, cudaStream_t st)
{
    StageScope sc(ST_PACK, st);
    const int n = g.P * g.nt;
    const size_t sm = sizeof(int) * n;
    cudaMemsetAsync(scratch, 0, sizeof(int) * g.B * n, st);
    if (skipped) cudaMemsetAsync(skipped, 0, sizeof(long long) * 2, st);
    const int64_t chunk = (int64_t)kPackThreads * kPackRows;
    dim3 grid((unsigned)((g.M + chunk - 1) / chunk), (unsigned)g.B);
    count_launch(g.M > 0 ? 3 : 1);
    if (g.M > 0) pack_count_kernel<<<grid, kPackThreads, sm, st>>>(events, g, scratch, skipped);
    pack_scan_kernel<<<(unsigned)g.B, 1024, 0, st>>>(g, scratch, seg_start);
    if (g.M > 0) pack_scatter_kernel<<<grid, kPackThreads, sm, st>>>(events, g, scratch, records);
    return check_launch();
}

// ---------------------------------------------------------------------------------------------
// shared pieces of the tile kernels
// ---------------------------------------------------------------------------------------------
struct Seg {
    int a, e;          // record range of this CTA inside its sample
    int ty, tx, grp;
};

__device__ __forceinline__ Seg cta_segment(const Geom &g, const int *__restrict__ seg_start, int split)
{
    Seg s;
    const int tile = blockIdx.x / split, part = blockIdx.x - tile * split;
    s.grp = blockIdx.y;
    s.ty = tile / g.ntx;
    s.tx = tile - s.ty * g.ntx;
    const int *row = seg_start + (int64_t)blockIdx.z * (g.P * g.nt + 1) + s.grp * g.nt + tile;
    const int a = __ldg(row), e = __ldg(row + 1);
    const int chunk = (e - a + split - 1) / split;
    s.a = min(a + part * chunk, e);
    s.e = min(s.a + chunk, e);
    return s;
}

// Window origin for reference time r: the tile's pixel box grown by the range of the flows stored
// in the tile's LUT block, sampled on a 3 x 3 lattice of its cells in every bin (the LUT is a
// K-neighbour mean, i.e. smooth inside a 32-pixel tile).  Only a placement heuristic - votes
// outside the window take the global path - so forward and backward need not agree and NaN /
// huge flows are harmless.  Computed by warp 0, broadcast through shared memory.
__device__ __forceinline__ void window_origin(const Geom &g, const float *__restrict__ lut, int64_t b,
                                              const Seg &s, int r, int *s_org, int *oy, int *ox)
{
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float mny = INFINITY, mxy = -INFINITY, mnx = INFINITY, mxx = -INFINITY;
        const int y0 = s.ty * g.ct, x0 = s.tx * g.ct;
        const int y2 = min(y0 + g.ct, g.Hq) - 1, x2 = min(x0 + g.ct, g.Wq) - 1;
        const int ys[3] = {y0, (y0 + y2) >> 1, y2}, xs[3] = {x0, (x0 + x2) >> 1, x2};
        for (int bin = lane; bin < g.nb; bin += 32) {
            const float2 *base = reinterpret_cast<const float2 *>(lut) + ((b * g.nb + bin) * (int64_t)g.q) * g.R + r;
#pragma unroll
            for (int u = 0; u < 9; ++u) {
                const float2 f = __ldg(base + (int64_t)(ys[u / 3] * g.Wq + xs[u % 3]) * g.R);
                mny = fminf(mny, f.x); mxy = fmaxf(mxy, f.x);
                mnx = fminf(mnx, f.y); mxx = fmaxf(mxx, f.y);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
            mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
            mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
            mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        }
        if (lane == 0) {
            const int tp = g.ct * g.s;
            auto place = [&](int t0, float mn, float mx) {
                if (!(mn > -1e6f && mx < 1e6f)) return t0 - (kWin - tp) / 2;
                const int lo = t0 + (int)floorf(mn) - 2, hi = t0 + tp + (int)ceilf(mx) + 3;
                return hi - lo <= kWin ? lo : (lo + hi) / 2 - kWin / 2;
            };
            s_org[0] = place(s.ty * tp, mny, mxy);
            s_org[1] = place(s.tx * tp, mnx, mxx) & ~3;     // quads of the flush stay 16-byte aligned
        }
    }
    __syncthreads();
    *oy = s_org[0];
    *ox = s_org[1];
}

struct PackedEvent {
    EventRow e;
    int bin, iy, ix;
};

__device__ __forceinline__ PackedEvent unpack(const float4 rec)
{
    PackedEvent p;
    const unsigned meta = __float_as_uint(rec.w);
    p.bin = (int)(meta >> 24);
    p.iy = (int)((meta >> 12) & 0xfffu);
    p.ix = (int)(meta & 0xfffu);
    p.e.y = rec.x; p.e.x = rec.y; p.e.t = rec.z;
    p.e.p = 0.0f; p.e.bin = (float)p.bin; p.e.valid = 1.0f;
    return p;
}

// corners with (row, column) kept separate, plus the per-corner in-image flags of
// event_image_converter.py:362-377
struct Corners2 {
    int y, x;            // integer (y1, x1); only meaningful when `finite`
    bool finite;         // y1, x1 within int range of the image neighbourhood
    bool y0ok, y1ok, x0ok, x1ok;
    float fy, fx;
};

__device__ __forceinline__ Corners2 vote_corners2(float wy, float wx, int H, int W)
{
    Corners2 c;
    const float y1 = floorf(__fadd_rn(wy, kVoteEps));
    const float x1 = floorf(__fadd_rn(wx, kVoteEps));
    c.fy = __fsub_rn(wy, y1);
    c.fx = __fsub_rn(wx, x1);
    c.y0ok = (y1 >= 0.0f) && (y1 < (float)H);
    c.y1ok = (y1 >= -1.0f) && (y1 < (float)(H - 1));
    c.x0ok = (x1 >= 0.0f) && (x1 < (float)W);
    c.x1ok = (x1 >= -1.0f) && (x1 < (float)(W - 1));
    c.finite = (c.y0ok || c.y1ok) && (c.x0ok || c.x1ok);
    c.y = c.finite ? (int)y1 : 0;
    c.x = c.finite ? (int)x1 : 0;
    return c;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <bool DET>
__global__ void __launch_bounds__(kTileThreads, 6)
event_forward_tile_kernel(const float4 *__restrict__ records, const int *__restrict__ seg_start,
                          const float *__restrict__ times, Geom g, int split,
                          const float *__restrict__ lut, float *__restrict__ raw,
                          long long *__restrict__ raw_i64)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned *s_lo = reinterpret_cast<unsigned *>(smem_raw);          // [kWin * kWin] low words
    unsigned *s_hi = s_lo + kWin * kWin;                              // [kWin * kWin] high words (DET only)
    __shared__ int s_org[2];
    __shared__ int s_rowflag[kWin];                                   // float mode: rows that received a vote

    const Seg sg = cta_segment(g, seg_start, split);
    if (sg.a >= sg.e) return;
    const int64_t b = blockIdx.z;
    const int tid = threadIdx.x;
    const int64_t HW = (int64_t)g.H * g.W;
    const float4 *recs = records + b * g.M;
    const bool use_win = sg.e - sg.a >= 64;              // tiny slices: global reds are cheaper
    // rows of the window can go out as bulk reduce-adds when every row start is 16-byte aligned
    const bool bulk_ok = !DET && (g.W & 3) == 0 && ((reinterpret_cast<uintptr_t>(raw) & 15u) == 0);

    const int lut_base = (int)b * g.nb;
    for (int r = 0; r < g.R; ++r) {
        int oy = 0, ox = 0;
        if (use_win) {
            for (int i = tid; i < (DET ? kWin * kWin / 2 : kWin * kWin / 4); i += kTileThreads)
                reinterpret_cast<uint4 *>(s_lo)[i] = make_uint4(0u, 0u, 0u, 0u);       // both arrays when DET
            if (tid < kWin) s_rowflag[tid] = 0;
            window_origin(g, lut, b, sg, r, s_org, &oy, &ox);       // ends with a barrier
        }
        // window entirely inside the image: every corner that is in the window is in bounds
        const bool interior = use_win && oy >= 0 && oy + kWin <= g.H && ox >= 0 && ox + kWin <= g.W;
        const float wy0 = use_win ? (float)oy : 1e30f, wy1 = (float)(oy + kWin - 1);
        const float wx0 = (float)ox, wx1 = (float)(ox + kWin - 1);
        const float tref = __ldg(times + r);
        const int64_t base = ((b * g.R + r) * g.P + sg.grp) * HW;
        float *img = raw + base;
        for (int i = sg.a + tid; i < sg.e; i += kTileThreads) {
            const PackedEvent pe = unpack(ld_stream_f4(recs + i));
            if (pe.bin >= g.nb || pe.iy >= g.Hq || pe.ix >= g.Wq) continue;      // corrupt record
            const int cell = ((lut_base + pe.bin) * g.Hq + pe.iy) * g.Wq + pe.ix;
            const float2 f = __ldg(reinterpret_cast<const float2 *>(lut) + cell * g.R + r);
            const float wy = __fadd_rn(f.x, pe.e.y), wx = __fadd_rn(f.y, pe.e.x);     // focus.py:191
            const float w = event_weight(pe.e, wy, wx, tref, g);
            if (w == 0.0f) continue;
            const float y1 = floorf(__fadd_rn(wy, kVoteEps)), x1 = floorf(__fadd_rn(wx, kVoteEps));
            const float fy = __fsub_rn(wy, y1), fx = __fsub_rn(wx, x1);
            const float oyw = __fsub_rn(1.0f, fy), oxw = __fsub_rn(1.0f, fx);
            float v[4];
            v[0] = __fmul_rn(__fmul_rn(oyw, oxw), w);        // event_image_converter.py:382-385
            v[1] = __fmul_rn(__fmul_rn(fy, oxw), w);
            v[2] = __fmul_rn(__fmul_rn(oyw, fx), w);
            v[3] = __fmul_rn(__fmul_rn(fy, fx), w);
            const bool in_win = y1 >= wy0 && y1 < wy1 && x1 >= wx0 && x1 < wx1;
            if (in_win && interior) {
                const int p = ((int)y1 - oy) * kWin + ((int)x1 - ox);
                float *gp = img + (int64_t)(int)y1 * g.W + (int)x1;      // the same pixel in HBM (carries)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int q = p + (k & 1) * kWin + (k >> 1);
                    if (DET) smem_add_fix(s_lo + q, s_hi + q, to_fix(v[k]));
                    else smem_add_vote(s_lo + q, v[k], gp + (k & 1) * g.W + (k >> 1));
                }
                continue;
            }
            const Corners2 c = vote_corners2(wy, wx, g.H, g.W);
            if (!c.finite) continue;
            const bool ok[4] = {c.y0ok && c.x0ok, c.y1ok && c.x0ok, c.y0ok && c.x1ok, c.y1ok && c.x1ok};
            if (in_win) {
                const int p = (c.y - oy) * kWin + (c.x - ox);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!ok[k]) continue;
                    const int q = p + (k & 1) * kWin + (k >> 1);
                    if (DET) smem_add_fix(s_lo + q, s_hi + q, to_fix(v[k]));
                    else smem_add_vote(s_lo + q, v[k], img + (int64_t)(c.y + (k & 1)) * g.W + c.x + (k >> 1));
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!ok[k]) continue;
                    const int64_t idx = base + (int64_t)(c.y + (k & 1)) * g.W + c.x + (k >> 1);
                    if (DET)
                        atomicAdd(reinterpret_cast<unsigned long long *>(raw_i64 + idx),
                                  (unsigned long long)to_fix(v[k]));
                    else
                        atomicAdd(raw + idx, v[k]);
                }
            }
        }
        if (!use_win) continue;
        __syncthreads();
        // flush the non-zero part of the window
        if (DET) {
            for (int i = tid; i < kWin * kWin; i += kTileThreads) {
                const long long v = smem_get_fix(s_lo + i, s_hi + i);
                const int gy = oy + i / kWin, gx = ox + i % kWin;
                if (v != 0 && gy >= 0 && gy < g.H && gx >= 0 && gx < g.W)
                    atomicAdd(reinterpret_cast<unsigned long long *>(raw_i64 + base + (int64_t)gy * g.W + gx),
                              (unsigned long long)v);
            }
        } else {
            // words -> float32 in place (one rounding of the exact tile sum), note the rows in use
            float *s_f = reinterpret_cast<float *>(s_lo);
            for (int i = tid; i < kWin * kWin / 4; i += kTileThreads) {       // 16 threads per window row
                const uint4 u = reinterpret_cast<const uint4 *>(s_lo)[i];
                const bool any = (u.x | u.y | u.z | u.w) != 0u;
                reinterpret_cast<float4 *>(s_f)[i] =
                    make_float4((float)u.x * (1.0f / kFloatFixScale), (float)u.y * (1.0f / kFloatFixScale),
                                (float)u.z * (1.0f / kFloatFixScale), (float)u.w * (1.0f / kFloatFixScale));
                const unsigned m = __ballot_sync(0xffffffffu, any);
                if ((tid & 15) == 0 && ((m >> (tid & 16)) & 0xffffu)) s_rowflag[i / (kWin / 4)] = 1;
            }
            if (bulk_ok) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> async proxy
                __syncthreads();
                if (tid < kWin && s_rowflag[tid]) {
                    const int gy = oy + tid;
                    const int c0 = max(0, -ox), c1 = min(kWin, g.W - ox);          // multiples of 4
                    if (gy >= 0 && gy < g.H && c1 > c0)
                        bulk_reduce_add_f32(img + (int64_t)gy * g.W + ox + c0, s_f + tid * kWin + c0,
                                            (unsigned)(c1 - c0) * 4u);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");        // window reusable / CTA may exit
            } else {
                __syncthreads();
                for (int i = tid; i < kWin * kWin / 4; i += kTileThreads) {
                    const float4 vv = reinterpret_cast<const float4 *>(s_f)[i];
                    if (vv.x == 0.0f && vv.y == 0.0f && vv.z == 0.0f && vv.w == 0.0f) continue;
                    const int gy = oy + i / (kWin / 4), gx = ox + 4 * (i % (kWin / 4));
                    if (gy < 0 || gy >= g.H) continue;
                    const int64_t idx = base + (int64_t)gy * g.W + gx;
                    const float q4[4] = {vv.x, vv.y, vv.z, vv.w};
                    if (gx >= 0 && gx + 3 < g.W && (idx & 3) == 0) {
                        red_add_f32x4(raw + idx, q4[0], q4[1], q4[2], q4[3]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (q4[k] != 0.0f && gx + k >= 0 && gx + k < g.W) atomicAdd(raw + idx + k, q4[k]);
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
template <bool DET>
__global__ void __launch_bounds__(kTileThreads, 6)
event_backward_tile_kernel(const float4 *__restrict__ records, const int *__restrict__ seg_start,
                           const float *__restrict__ times, Geom g, int split, int smem_acc,
                           const float *__restrict__ lut, const float *__restrict__ dimg,
                           const Header *__restrict__ hdr, const float *__restrict__ grad_loss,
                           float *__restrict__ dlut, long long *__restrict__ dlut_i64)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_D = reinterpret_cast<float *>(smem_raw);                       // [kWin * kWin]
    // dLUT block of the tile, [nb, ct, ct, 2] in 2^-32 fixed point: low words, then high words
    unsigned *s_alo = reinterpret_cast<unsigned *>(smem_raw + sizeof(float) * kWin * kWin);
    __shared__ int s_org[2];
    __shared__ __align__(8) unsigned long long s_bar;       // completion of the bulk copies of the window
    unsigned bar_phase = 0u;

    const Seg sg = cta_segment(g, seg_start, split);
    if (sg.a >= sg.e) return;
    const int64_t b = blockIdx.z;
    const int tid = threadIdx.x;
    const int64_t HW = (int64_t)g.H * g.W;
    const float4 *recs = records + b * g.M;
    const bool use_win = sg.e - sg.a >= 64;
    const bool use_acc = use_win && smem_acc;
    const int cc = g.ct * g.ct, nacc = cc * g.nb;
    unsigned *s_ahi = s_alo + 2 * nacc;
    float coef = 1.0f;
    if (!DET) {
        const float val = hdr->val;
        const float N = (float)((double)g.B * g.R * g.P * (double)HW);
        coef = __ldg(grad_loss) * (-(1.0f / (val * val))) / N;
    }

    if (tid == 0) mbar_init(&s_bar, 1);                      // visible to all after window_origin's barrier
    const int lut_base = (int)b * g.nb;
    for (int r = 0; r < g.R; ++r) {
        const float *D = dimg + ((b * g.R + r) * g.P + sg.grp) * HW;
        int oy = 0, ox = 0;
        if (use_win) {
            if (use_acc)
                for (int i = tid; i < nacc * 4; i += kTileThreads) s_alo[i] = 0u;       // both arrays
            window_origin(g, lut, b, sg, r, s_org, &oy, &ox);
            // stage dL/dIWE of the window (zero outside the image)
            const bool aligned = (g.W & 3) == 0 && ((((b * g.R + r) * g.P + sg.grp) * HW) & 3) == 0 &&
                                 (reinterpret_cast<uintptr_t>(dimg) & 15u) == 0;
            if (aligned && ox >= 0 && ox + kWin <= g.W) {
                // every window row inside the image is one contiguous, 16-byte aligned run of 256 B:
                // one bulk async copy per row (TMA engine, completion on the mbarrier); rows above /
                // below the image are zero-filled by the threads
                const int r_lo = min(max(0, -oy), kWin), r_hi = max(min(kWin, g.H - oy), r_lo);
                if (tid == 0) mbar_expect_tx(&s_bar, (unsigned)(r_hi - r_lo) * kWin * 4u);
                if (tid >= r_lo && tid < r_hi)
                    bulk_g2s(s_D + tid * kWin, D + (int64_t)(oy + tid) * g.W + ox, kWin * 4u, &s_bar);
                for (int i = tid; i < kWin * kWin / 4; i += kTileThreads) {
                    const int row = i / (kWin / 4);
                    if (row < r_lo || row >= r_hi) reinterpret_cast<float4 *>(s_D)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                mbar_wait(&s_bar, bar_phase);
                bar_phase ^= 1u;
            } else if (aligned) {
                for (int i = tid; i < kWin * kWin / 4; i += kTileThreads) {
                    const int gy = oy + i / (kWin / 4), gx = ox + 4 * (i % (kWin / 4));   // ox % 4 == 0
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W)
                        v = __ldg(reinterpret_cast<const float4 *>(D + (int64_t)gy * g.W + gx));
                    reinterpret_cast<float4 *>(s_D)[i] = v;
                }
            } else {
                for (int i = tid; i < kWin * kWin; i += kTileThreads) {
                    const int gy = oy + i / kWin, gx = ox + i % kWin;
                    s_D[i] = (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) ? __ldg(D + (int64_t)gy * g.W + gx) : 0.0f;
                }
            }
            __syncthreads();
        }
        const bool interior = use_win && oy >= 0 && oy + kWin <= g.H && ox >= 0 && ox + kWin <= g.W;
        const float wy0 = use_win ? (float)oy : 1e30f, wy1 = (float)(oy + kWin - 1);
        const float wx0 = (float)ox, wx1 = (float)(ox + kWin - 1);
        const float tref = __ldg(times + r);
        for (int i = sg.a + tid; i < sg.e; i += kTileThreads) {
            const PackedEvent pe = unpack(ld_stream_f4(recs + i));
            if (pe.bin >= g.nb || pe.iy >= g.Hq || pe.ix >= g.Wq) continue;      // corrupt record
            const int cell = ((lut_base + pe.bin) * g.Hq + pe.iy) * g.Wq + pe.ix;
            const float2 f = __ldg(reinterpret_cast<const float2 *>(lut) + cell * g.R + r);
            const float wy = __fadd_rn(f.x, pe.e.y), wx = __fadd_rn(f.y, pe.e.x);
            const float w = event_weight(pe.e, wy, wx, tref, g);
            if (w == 0.0f) continue;
            const float y1 = floorf(__fadd_rn(wy, kVoteEps)), x1 = floorf(__fadd_rn(wx, kVoteEps));
            const float fy = __fsub_rn(wy, y1), fx = __fsub_rn(wx, x1);
            const bool in_win = y1 >= wy0 && y1 < wy1 && x1 >= wx0 && x1 < wx1;
            float d00, d10, d01, d11;
            if (in_win && interior) {
                const float *p = s_D + ((int)y1 - oy) * kWin + ((int)x1 - ox);
                d00 = p[0]; d10 = p[kWin]; d01 = p[1]; d11 = p[kWin + 1];
            } else {
                const Corners2 c = vote_corners2(wy, wx, g.H, g.W);
                if (!c.finite) continue;
                if (in_win) {
                    const float *p = s_D + (c.y - oy) * kWin + (c.x - ox);      // zero outside the image
                    d00 = (c.y0ok && c.x0ok) ? p[0] : 0.0f;
                    d10 = (c.y1ok && c.x0ok) ? p[kWin] : 0.0f;
                    d01 = (c.y0ok && c.x1ok) ? p[1] : 0.0f;
                    d11 = (c.y1ok && c.x1ok) ? p[kWin + 1] : 0.0f;
                } else {
                    const float *p = D + (int64_t)c.y * g.W + c.x;
                    d00 = (c.y0ok && c.x0ok) ? __ldg(p) : 0.0f;
                    d10 = (c.y1ok && c.x0ok) ? __ldg(p + g.W) : 0.0f;
                    d01 = (c.y0ok && c.x1ok) ? __ldg(p + 1) : 0.0f;
                    d11 = (c.y1ok && c.x1ok) ? __ldg(p + g.W + 1) : 0.0f;
                }
            }
            const float oyw = 1.0f - fy, oxw = 1.0f - fx;
            const float gy = w * (oxw * (d10 - d00) + fx * (d11 - d01));
            const float gx = w * (oyw * (d01 - d00) + fy * (d11 - d10));
            const int cy = pe.iy - sg.ty * g.ct, cx = pe.ix - sg.tx * g.ct;
            if (use_acc && (unsigned)cy < (unsigned)g.ct && (unsigned)cx < (unsigned)g.ct) {
                const int q = 2 * ((pe.bin * g.ct + cy) * g.ct + cx);
                smem_add_fix(s_alo + q, s_ahi + q, to_fix(gy));
                smem_add_fix(s_alo + q + 1, s_ahi + q + 1, to_fix(gx));
            } else if (DET) {
                unsigned long long *dst = reinterpret_cast<unsigned long long *>(dlut_i64) + ((int64_t)cell * g.R + r) * 2;
                atomicAdd(dst, (unsigned long long)to_fix(gy));
                atomicAdd(dst + 1, (unsigned long long)to_fix(gx));
            } else {
                red_add_f32x2(dlut + ((int64_t)cell * g.R + r) * 2, coef * gy, coef * gx);
            }
        }
        if (!use_win) continue;
        __syncthreads();
        if (use_acc) {
            for (int i = tid; i < nacc; i += kTileThreads) {
                const int bin = i / cc, c2 = i - bin * cc;
                const int iy = sg.ty * g.ct + c2 / g.ct, ix = sg.tx * g.ct + c2 % g.ct;
                if (iy >= g.Hq || ix >= g.Wq) continue;
                const int64_t cell = ((b * g.nb + bin) * g.Hq + iy) * g.Wq + ix;
                const long long a0 = smem_get_fix(s_alo + 2 * i, s_ahi + 2 * i);
                const long long a1 = smem_get_fix(s_alo + 2 * i + 1, s_ahi + 2 * i + 1);
                if (a0 == 0 && a1 == 0) continue;
                if (DET) {
                    unsigned long long *dst = reinterpret_cast<unsigned long long *>(dlut_i64) + (cell * g.R + r) * 2;
                    if (a0 != 0) atomicAdd(dst, (unsigned long long)a0);
                    if (a1 != 0) atomicAdd(dst + 1, (unsigned long long)a1);
                } else {
                    red_add_f32x2(dlut + (cell * g.R + r) * 2,
                                  coef * (__ll2float_rn(a0) * (1.0f / 4294967296.0f)),
                                  coef * (__ll2float_rn(a1) * (1.0f / 4294967296.0f)));
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static int pick_split(const Geom &g)
{
    // slices of ~8k events: enough CTAs for every SM when a window holds tens of millions of events
    const int64_t per_seg = g.M / ((int64_t)g.P * g.nt) + 1;
    int64_t s = (per_seg + 8191) / 8192;
    return (int)(s < 1 ? 1 : (s > 32 ? 32 : s));
}

template <class K>
static int set_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024 &&
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
        return CMAX_ERR_CUDA;
    return CMAX_OK;
}

int launch_event_forward_packed(const Geom &g, const Layout &L, const float4 *records,
                                const int *seg_start, const float *times, char *ws, cudaStream_t st)
{
    const int64_t count = g.B * g.R * g.P * (int64_t)g.H * g.W;
    float *raw = reinterpret_cast<float *>(ws + L.raw);
    long long *raw_i64 = reinterpret_cast<long long *>(ws + L.raw_i64);
    const float *lut = reinterpret_cast<const float *>(ws + L.lut);
    StageScope sc(ST_EVENT_FWD, st);
    count_launch((g.M > 0 ? 1 : 0) + (g.det ? 1 : 0));
    if (g.det)
        cudaMemsetAsync(raw_i64, 0, sizeof(long long) * count, st);
    else
        cudaMemsetAsync(raw, 0, sizeof(float) * count, st);
    if (g.M > 0) {
        const int split = pick_split(g);
        dim3 grid((unsigned)(g.nt * split), (unsigned)g.P, (unsigned)g.B);
        const size_t sm = (g.det ? 2 : 1) * sizeof(unsigned) * kWin * kWin;
        if (g.det)
            event_forward_tile_kernel<true><<<grid, kTileThreads, sm, st>>>(records, seg_start, times, g,
                                                                            split, lut, raw, raw_i64);
        else
            event_forward_tile_kernel<false><<<grid, kTileThreads, sm, st>>>(records, seg_start, times, g,
                                                                             split, lut, raw, raw_i64);
    }
    if (g.det) return launch_fix_to_float(raw_i64, raw, count, st);
    return check_launch();
}

int launch_event_backward_packed(const Geom &g, const Layout &L, const float4 *records,
                                 const int *seg_start, const float *times, const float *grad_loss,
                                 char *ws, cudaStream_t st)
{
    const int64_t count = g.S * g.q * g.R * 2;
    const Header *hdr = reinterpret_cast<const Header *>(ws + L.header);
    const float *lut = reinterpret_cast<const float *>(ws + L.lut);
    const float *dimg = reinterpret_cast<const float *>(ws + L.dimg);
    float *dlut = reinterpret_cast<float *>(ws + L.dlut);
    long long *dlut_i64 = reinterpret_cast<long long *>(ws + L.dlut_i64);
    StageScope sc(ST_EVENT_BWD, st);
    count_launch((g.M > 0 ? 1 : 0) + (g.det ? 1 : 0));
    if (g.det) cudaMemsetAsync(dlut_i64, 0, sizeof(long long) * count, st);
    if (g.M > 0) {
        const int split = pick_split(g);
        dim3 grid((unsigned)(g.nt * split), (unsigned)g.P, (unsigned)g.B);
        const size_t acc = (size_t)g.ct * g.ct * g.nb * 2 * 2 * sizeof(unsigned);
        const int smem_acc = acc <= (size_t)kMaxSmemAcc ? 1 : 0;
        const size_t sm = sizeof(float) * kWin * kWin + (smem_acc ? acc : 0);
        int rc;
        if (g.det) {
            if ((rc = set_smem(event_backward_tile_kernel<true>, sm))) return rc;
            event_backward_tile_kernel<true><<<grid, kTileThreads, sm, st>>>(
                records, seg_start, times, g, split, smem_acc, lut, dimg, hdr, grad_loss, dlut, dlut_i64);
        } else {
            if ((rc = set_smem(event_backward_tile_kernel<false>, sm))) return rc;
            event_backward_tile_kernel<false><<<grid, kTileThreads, sm, st>>>(
                records, seg_start, times, g, split, smem_acc, lut, dimg, hdr, grad_loss, dlut, dlut_i64);
        }
    }
    if (g.det) return launch_dlut_finalize(g, L, grad_loss, ws, st);
    return check_launch();
}

}  // namespace cmax
