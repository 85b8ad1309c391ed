// host_pack.cpp - HOST-side builder of the packed, tile-binned event layout (include/cmax_b200.h),
// for DataLoader workers / the collate function.  Replaces the producer of the `events` tensor
// (upstream src/loader/dsec/loader.py:141-182,360-415) for the loss path; plain C++ + OpenMP, no
// CUDA call.  Produces exactly what motionpriorcmax_b200.io.pack_events_host (torch CPU ops, stable
// sort) produces: a counting sort by (polarity group, source tile) that keeps the row order inside
// every segment.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/cmax_b200.h"

namespace {

// torch's `//` on float32 (c10::div_floor_floating), the arithmetic of focus.py:186-187
inline float floordiv_f32_generic(float a, float b)
{
    const float mod = fmodf(a, b);
    float div = (a - mod) / b;
    if (mod != 0.0f && ((b < 0.0f) != (mod < 0.0f))) div -= 1.0f;
    float fd;
    if (div != 0.0f) {
        fd = floorf(div);
        if (div - fd > 0.5f) fd += 1.0f;
    } else {
        fd = copysignf(0.0f, a / b);
    }
    return fd;
}

// Same fast path as the device code (cmax_common.cuh): for a power-of-two divisor fmod is exact,
// a - mod an exact multiple of b, and the recipe reduces to floor(a / b); a * (1 / b) is exact as
// long as it does not underflow.  Two fmodf calls per row were most of the packer's time.
inline float floordiv_f32(float a, float b)
{
    uint32_t bb;
    memcpy(&bb, &b, 4);
    const float aa = fabsf(a);
    if (b > 0.0f && (bb & 0x007fffffu) == 0u && aa >= 1e-30f && aa < 1e30f) return floorf(a * (1.0f / b));
    return floordiv_f32_generic(a, b);
}

struct HostLayout {
    int s, nb, Hq, Wq, ct, nty, ntx, nt, G;
    int ct_shift;      // log2(ct) when ct is a power of two (every shipped config), else -1
};

bool host_layout(const CmaxConfig *c, HostLayout *L)
{
    int32_t out[4];
    if (cmax_pack_layout(c, out) != CMAX_OK) return false;
    L->s = c->lut_superpixel_size;
    L->nb = c->num_bins;
    L->Hq = (c->height + L->s - 1) / L->s;
    L->Wq = (c->width + L->s - 1) / L->s;
    L->ct = out[0];
    L->nty = out[1];
    L->ntx = out[2];
    L->nt = L->nty * L->ntx;
    L->G = out[3];
    L->ct_shift = -1;
    for (int sh = 0; sh < 16; ++sh)
        if ((1 << sh) == L->ct) L->ct_shift = sh;
    return true;
}

// segment key and meta word of one row; -1: dropped (padding), -2: dropped (cell outside the table)
inline int row_key(const float *row, int64_t m, int64_t npos, const HostLayout &L, uint32_t *meta)
{
    const float valid = row[5];
    if (valid == 0.0f) return -1;
    const float fs = (float)L.s;
    const float fy = floordiv_f32(row[0], fs), fx = floordiv_f32(row[1], fs);
    const float ft = truncf(row[4]);
    if (!(ft >= 0.0f && ft < (float)L.nb && fy >= 0.0f && fy < (float)L.Hq && fx >= 0.0f && fx < (float)L.Wq))
        return -2;
    const int it = (int)ft, iy = (int)fy, ix = (int)fx;
    const int grp = (L.G == 2 && m >= npos) ? 1 : 0;
    *meta = ((uint32_t)it << 24) | ((uint32_t)iy << 12) | (uint32_t)ix;
    if (L.ct_shift >= 0) return grp * L.nt + (iy >> L.ct_shift) * L.ntx + (ix >> L.ct_shift);
    return grp * L.nt + (iy / L.ct) * L.ntx + ix / L.ct;
}

}  // namespace

extern "C" int cmax_pack_events_host(const CmaxConfig *cfg, const float *events_host, int64_t B,
                                     int64_t M, int64_t num_pos_events, float *records_host,
                                     int64_t records_stride, int32_t *seg_start_host,
                                     int64_t *skipped_host)
{
    HostLayout L;
    if (!cfg) return CMAX_ERR_BAD_CONFIG;
    if (!host_layout(cfg, &L)) return CMAX_ERR_UNSUPPORTED;
    if (B < 0 || M < 0 || (!events_host && B * M > 0) || !seg_start_host) return CMAX_ERR_BAD_SHAPE;
    if (L.G == 2 && (num_pos_events < 0 || num_pos_events > M)) return CMAX_ERR_BAD_SHAPE;
    if (M > (int64_t)INT32_MAX) return CMAX_ERR_UNSUPPORTED;
    const int nkeys = L.G * L.nt;
    int64_t dropped = 0, odd = 0;
    int rc = CMAX_OK;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : dropped, odd)
    for (int64_t b = 0; b < B; ++b) {
        const float *ev = events_host + b * M * 6;
        int32_t *seg = seg_start_host + b * (int64_t)(nkeys + 1);
        std::vector<int32_t> key((size_t)M);
        std::vector<uint32_t> meta((size_t)M);
        std::vector<int32_t> cursor((size_t)nkeys + 1, 0);
        for (int64_t m = 0; m < M; ++m) {
            uint32_t mw = 0;
            const int k = row_key(ev + m * 6, m, num_pos_events, L, &mw);
            key[(size_t)m] = k;
            meta[(size_t)m] = mw;
            if (k >= 0) {
                ++cursor[(size_t)k + 1];
                if (ev[m * 6 + 5] != 1.0f) ++odd;
            } else if (k == -2) {
                ++dropped;
                if (ev[m * 6 + 5] != 1.0f) ++odd;
            }
        }
        for (int k = 0; k < nkeys; ++k) cursor[(size_t)k + 1] += cursor[(size_t)k];
        memcpy(seg, cursor.data(), sizeof(int32_t) * (size_t)(nkeys + 1));
        if (records_host) {
            if (cursor[(size_t)nkeys] > records_stride) {
#pragma omp critical
                rc = CMAX_ERR_BAD_SHAPE;              // records_stride too small for this window
                continue;
            }
            float *rec = records_host + b * records_stride * 4;
            for (int64_t m = 0; m < M; ++m) {
                const int k = key[(size_t)m];
                if (k < 0) continue;
                float *dst = rec + (int64_t)cursor[(size_t)k]++ * 4;
                dst[0] = ev[m * 6 + 0];
                dst[1] = ev[m * 6 + 1];
                dst[2] = ev[m * 6 + 2];
                memcpy(dst + 3, &meta[(size_t)m], 4);
            }
        }
    }
    if (skipped_host) {
        skipped_host[0] = dropped;
        skipped_host[1] = odd;
    }
    return rc;
}


// Compact wire layout (12 bytes per valid event instead of 16 / 24): what crosses PCIe every step.
//   coords     [T, 3] float32 (y, x, t), the windows of the batch back to back (no padding):
//              window b owns rows sample_off[b] .. sample_off[b + 1];
//   fine_start [B, G * NT * nb + 1] int32: per window, prefix offsets (relative to the window's
//              first row) of the runs ordered by (polarity group, source tile, time bin) - the
//              time bin (focus.py:185: it = int(bin)) is implied by the run, the LUT cell by (y, x).
// cmax_expand_compact (device) turns it into the 16-byte records + seg_start of the packed layout.
// Same stable counting sort as above with the finer key; rows keep their order inside a run.
extern "C" int cmax_pack_events_host_compact(const CmaxConfig *cfg, const float *events_host, int64_t B,
                                             int64_t M, int64_t num_pos_events, float *coords_host,
                                             int64_t coords_capacity, int32_t *fine_start_host,
                                             int64_t *sample_off_host, int64_t *skipped_host)
{
    HostLayout L;
    if (!cfg) return CMAX_ERR_BAD_CONFIG;
    if (!host_layout(cfg, &L)) return CMAX_ERR_UNSUPPORTED;
    if (B < 0 || M < 0 || (!events_host && B * M > 0) || !fine_start_host || !sample_off_host) return CMAX_ERR_BAD_SHAPE;
    if (L.G == 2 && (num_pos_events < 0 || num_pos_events > M)) return CMAX_ERR_BAD_SHAPE;
    if (M > (int64_t)INT32_MAX) return CMAX_ERR_UNSUPPORTED;
    const int nkeys = L.G * L.nt * L.nb;
    int64_t dropped = 0, odd = 0;
    const bool fill = coords_host != nullptr;       // second call: sample_off_host is an input
    if (fill && sample_off_host[B] > coords_capacity) return CMAX_ERR_BAD_SHAPE;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : dropped, odd)
    for (int64_t b = 0; b < B; ++b) {
        const float *ev = events_host + b * M * 6;
        int32_t *seg = fine_start_host + b * (int64_t)(nkeys + 1);
        std::vector<int32_t> key((size_t)M);
        std::vector<int32_t> cursor((size_t)nkeys + 1, 0);
        for (int64_t m = 0; m < M; ++m) {
            uint32_t mw = 0;
            int k = row_key(ev + m * 6, m, num_pos_events, L, &mw);
            if (k >= 0) {
                k = k * L.nb + (int)(mw >> 24);
                ++cursor[(size_t)k + 1];
                if (ev[m * 6 + 5] != 1.0f) ++odd;
            } else if (k == -2) {
                ++dropped;
                if (ev[m * 6 + 5] != 1.0f) ++odd;
            }
            key[(size_t)m] = k;
        }
        for (int k = 0; k < nkeys; ++k) cursor[(size_t)k + 1] += cursor[(size_t)k];
        memcpy(seg, cursor.data(), sizeof(int32_t) * (size_t)(nkeys + 1));
        if (fill) {
            float *rec = coords_host + sample_off_host[b] * 3;
            for (int64_t m = 0; m < M; ++m) {
                const int k = key[(size_t)m];
                if (k < 0) continue;
                float *dst = rec + (int64_t)cursor[(size_t)k]++ * 3;
                dst[0] = ev[m * 6 + 0];
                dst[1] = ev[m * 6 + 1];
                dst[2] = ev[m * 6 + 2];
            }
        }
    }
    if (!fill) {                                     // first call: window offsets from the counts
        sample_off_host[0] = 0;
        for (int64_t b = 0; b < B; ++b)
            sample_off_host[b + 1] = sample_off_host[b] + fine_start_host[b * (int64_t)(nkeys + 1) + nkeys];
    }
    if (skipped_host) {
        skipped_host[0] = dropped;
        skipped_host[1] = odd;
    }
    return CMAX_OK;
}


// Bit-packed wire layout: the same runs as the compact layout, but every run stores its events as
// fixed-width records of bit-pattern DELTAS.  Inside one (group, tile, bin) run the three float32
// fields vary little: y and x stay inside a 32 x 32 pixel tile, t inside one time bin.  For
// non-negative floats the IEEE bit pattern is monotone, so a field is stored exactly as
//     bits(value) - min over the run of bits(value)        in   w = bit_length(max - min)   bits
// (w = 0 .. 32 per field, whatever the run needs: about 21 + 20 + 21 = 62 bits for a DSEC window
// instead of 96).  Lossless for ANY float32 input - negative values, NaN payloads and -0 merely
// make a run wide.
//   run_hdr  [B, F, 4] uint32: min bit patterns of y, x, t and  wy | wx << 8 | wt << 16
//   run_word [B, F + 1] int32: first 32-bit word of every run inside the window's bit stream
//            (runs start word aligned; run_word[b, F] = words of the window)
//   words    [total]  uint32: the windows' streams back to back; window b starts at word_off[b]
//            (int64 [B + 1]); the stream of a window ends with two zero words of slack so that the
//            decoder may always read three words per field window
//   fine_start as in the compact layout (event counts per run)
// words_host = NULL: only the tables (fine_start / run_hdr / run_word / word_off = the sizes).
// words_host != NULL: tables AND streams in the same call; CMAX_ERR_BAD_SHAPE (tables and word_off
// filled in) when words_capacity is too small - so either two calls with an exact buffer, or one
// call with a buffer of B * (3 * M + F + 2) words, which always suffices.
extern "C" int cmax_pack_events_host_bitpacked(const CmaxConfig *cfg, const float *events_host, int64_t B,
                                               int64_t M, int64_t num_pos_events, uint32_t *words_host,
                                               int64_t words_capacity, int32_t *fine_start_host,
                                               uint32_t *run_hdr_host, int32_t *run_word_host,
                                               int64_t *word_off_host, int64_t *skipped_host)
{
    HostLayout L;
    if (!cfg) return CMAX_ERR_BAD_CONFIG;
    if (!host_layout(cfg, &L)) return CMAX_ERR_UNSUPPORTED;
    if (B < 0 || M < 0 || (!events_host && B * M > 0) || !fine_start_host || !run_hdr_host || !run_word_host ||
        !word_off_host)
        return CMAX_ERR_BAD_SHAPE;
    if (L.G == 2 && (num_pos_events < 0 || num_pos_events > M)) return CMAX_ERR_BAD_SHAPE;
    if (M > (int64_t)INT32_MAX) return CMAX_ERR_UNSUPPORTED;
    const int nkeys = L.G * L.nt * L.nb;
    int64_t dropped = 0, odd = 0;
    const bool fill = words_host != nullptr;
    int rc = CMAX_OK;
    // phase 1 (parallel over the windows): run of every row, counts, per-run minima / widths, the
    // words every window needs; phase 2 after the prefix over the windows: the bit streams.
    std::vector<std::vector<int32_t>> keys((size_t)B);
    std::vector<std::vector<uint64_t>> bitpos((size_t)B);                 // next free bit of every run
    std::vector<int64_t> win_words((size_t)B, 0);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : dropped, odd)
    for (int64_t b = 0; b < B; ++b) {
        const float *ev = events_host + b * M * 6;
        int32_t *seg = fine_start_host + b * (int64_t)(nkeys + 1);
        uint32_t *hdr = run_hdr_host + b * (int64_t)nkeys * 4;
        int32_t *rw = run_word_host + b * (int64_t)(nkeys + 1);
        // (the run tables are a few hundred KB: they stay in cache while the rows stream by)
        std::vector<int32_t> &key = keys[(size_t)b];
        if (fill) key.resize((size_t)M);
        std::vector<int32_t> count((size_t)nkeys, 0);
        std::vector<uint32_t> mn((size_t)nkeys * 3, 0xffffffffu), mx((size_t)nkeys * 3, 0u);
        for (int64_t m = 0; m < M; ++m) {
            uint32_t mw = 0;
            const float *row = ev + m * 6;
            int k = row_key(row, m, num_pos_events, L, &mw);
            if (k >= 0) {
                k = k * L.nb + (int)(mw >> 24);
                ++count[(size_t)k];
                if (row[5] != 1.0f) ++odd;
                uint32_t v[3];
                memcpy(v, row, 12);
                for (int c = 0; c < 3; ++c) {
                    if (v[c] < mn[(size_t)k * 3 + c]) mn[(size_t)k * 3 + c] = v[c];
                    if (v[c] > mx[(size_t)k * 3 + c]) mx[(size_t)k * 3 + c] = v[c];
                }
            } else if (k == -2) {
                ++dropped;
                if (row[5] != 1.0f) ++odd;
            }
            if (fill) key[(size_t)m] = k;
        }
        auto bitlen = [](uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; };
        int64_t word = 0;
        int32_t rows_before = 0;
        if (fill) bitpos[(size_t)b].resize((size_t)nkeys);
        for (int k = 0; k < nkeys; ++k) {
            const int32_t n = count[(size_t)k];
            int wy = 0, wx = 0, wt = 0;
            uint32_t m0 = 0u, m1 = 0u, m2 = 0u;
            if (n > 0) {
                m0 = mn[(size_t)k * 3];
                m1 = mn[(size_t)k * 3 + 1];
                m2 = mn[(size_t)k * 3 + 2];
                wy = bitlen(mx[(size_t)k * 3] - m0);
                wx = bitlen(mx[(size_t)k * 3 + 1] - m1);
                wt = bitlen(mx[(size_t)k * 3 + 2] - m2);
            }
            seg[k] = rows_before;
            rows_before += n;
            hdr[k * 4 + 0] = m0;
            hdr[k * 4 + 1] = m1;
            hdr[k * 4 + 2] = m2;
            hdr[k * 4 + 3] = (uint32_t)wy | ((uint32_t)wx << 8) | ((uint32_t)wt << 16);
            rw[k] = (int32_t)word;
            if (fill) bitpos[(size_t)b][(size_t)k] = (uint64_t)word * 32u;
            word += ((int64_t)n * (wy + wx + wt) + 31) / 32;
        }
        seg[nkeys] = rows_before;
        rw[nkeys] = (int32_t)(word < (int64_t)INT32_MAX ? word : (int64_t)INT32_MAX);
        win_words[(size_t)b] = word;
    }
    // window offsets (two slack words each): an output of every call
    word_off_host[0] = 0;
    for (int64_t b = 0; b < B; ++b) {
        if (win_words[(size_t)b] + 2 > (int64_t)INT32_MAX) rc = CMAX_ERR_UNSUPPORTED;
        word_off_host[b + 1] = word_off_host[b] + win_words[(size_t)b] + 2;
    }
    if (rc == CMAX_OK && fill && word_off_host[B] > words_capacity) rc = CMAX_ERR_BAD_SHAPE;   // sizes are filled in
    if (rc == CMAX_OK && fill) {
#pragma omp parallel for schedule(dynamic, 1)
        for (int64_t b = 0; b < B; ++b) {
            // every kept row goes straight to the next free bits of its run (rows are visited in
            // their original order, so the order inside a run is the original one)
            const float *ev = events_host + b * M * 6;
            const uint32_t *hdr = run_hdr_host + b * (int64_t)nkeys * 4;
            const std::vector<int32_t> &key = keys[(size_t)b];
            std::vector<uint64_t> &bp = bitpos[(size_t)b];
            uint32_t *out = words_host + word_off_host[b];
            memset(out, 0, sizeof(uint32_t) * (size_t)(win_words[(size_t)b] + 2));
            for (int64_t m = 0; m < M; ++m) {
                const int k = key[(size_t)m];
                if (k < 0) continue;
                const uint32_t wv = hdr[k * 4 + 3];
                uint32_t v[3];
                memcpy(v, ev + m * 6, 12);
                uint64_t bit = bp[(size_t)k];
                for (int c = 0; c < 3; ++c) {
                    const int w = (int)((wv >> (8 * c)) & 255u);
                    if (w == 0) continue;
                    const uint64_t d = (uint64_t)(v[c] - hdr[k * 4 + c]);
                    const uint64_t wi = bit >> 5;
                    const int sh = (int)(bit & 31u);
                    out[wi] |= (uint32_t)(d << sh);
                    if (sh + w > 32) out[wi + 1] |= (uint32_t)(d >> (32 - sh));
                    bit += (uint64_t)w;
                }
                bp[(size_t)k] = bit;
            }
        }
    }
    if (skipped_host) {
        skipped_host[0] = dropped;
        skipped_host[1] = odd;
    }
    return rc;
}
