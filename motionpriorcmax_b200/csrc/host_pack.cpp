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
inline float floordiv_f32(float a, float b)
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

struct HostLayout {
    int s, nb, Hq, Wq, ct, nty, ntx, nt, G;
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
// Two calls like the compact packer: words_host = NULL fills fine_start / run_hdr / run_word /
// word_off (sizes), the second call writes the stream.
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
    if (fill && word_off_host[B] > words_capacity) return CMAX_ERR_BAD_SHAPE;
    int rc = CMAX_OK;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : dropped, odd)
    for (int64_t b = 0; b < B; ++b) {
        const float *ev = events_host + b * M * 6;
        int32_t *seg = fine_start_host + b * (int64_t)(nkeys + 1);
        uint32_t *hdr = run_hdr_host + b * (int64_t)nkeys * 4;
        int32_t *rw = run_word_host + b * (int64_t)(nkeys + 1);
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
        // stable gather of the kept rows into run order (bit patterns of y, x, t)
        const int32_t total = cursor[(size_t)nkeys];
        std::vector<uint32_t> rows((size_t)total * 3);
        for (int64_t m = 0; m < M; ++m) {
            const int k = key[(size_t)m];
            if (k < 0) continue;
            uint32_t *dst = rows.data() + (size_t)cursor[(size_t)k]++ * 3;
            memcpy(dst, ev + m * 6, 12);
        }
        // per run: minima, widths, first word
        auto bitlen = [](uint32_t v) { int n = 0; while (v) { ++n; v >>= 1; } return n; };
        int64_t word = 0;
        for (int k = 0; k < nkeys; ++k) {
            const int32_t a = seg[k], e = seg[k + 1];
            uint32_t mn[3] = {0u, 0u, 0u}, mx[3] = {0u, 0u, 0u};
            for (int32_t i = a; i < e; ++i)
                for (int c = 0; c < 3; ++c) {
                    const uint32_t v = rows[(size_t)i * 3 + c];
                    if (i == a || v < mn[c]) mn[c] = v;
                    if (i == a || v > mx[c]) mx[c] = v;
                }
            const int wy = bitlen(mx[0] - mn[0]), wx = bitlen(mx[1] - mn[1]), wt = bitlen(mx[2] - mn[2]);
            hdr[k * 4 + 0] = mn[0];
            hdr[k * 4 + 1] = mn[1];
            hdr[k * 4 + 2] = mn[2];
            hdr[k * 4 + 3] = (uint32_t)wy | ((uint32_t)wx << 8) | ((uint32_t)wt << 16);
            rw[k] = (int32_t)word;
            word += ((int64_t)(e - a) * (wy + wx + wt) + 31) / 32;
        }
        rw[nkeys] = (int32_t)word;
        if (word + 2 > (int64_t)INT32_MAX) {
#pragma omp critical
            rc = CMAX_ERR_UNSUPPORTED;
            continue;
        }
        if (fill) {
            uint32_t *out = words_host + word_off_host[b];
            memset(out, 0, sizeof(uint32_t) * (size_t)(word + 2));
            for (int k = 0; k < nkeys; ++k) {
                const int32_t a = seg[k], e = seg[k + 1];
                const uint32_t wv = hdr[k * 4 + 3];
                const int w3[3] = {(int)(wv & 255u), (int)((wv >> 8) & 255u), (int)((wv >> 16) & 255u)};
                uint64_t bit = (uint64_t)rw[k] * 32u;
                for (int32_t i = a; i < e; ++i)
                    for (int c = 0; c < 3; ++c) {
                        const int w = w3[c];
                        if (w == 0) continue;
                        const uint64_t v = (uint64_t)(rows[(size_t)i * 3 + c] - hdr[k * 4 + c]);
                        const uint64_t wi = bit >> 5;
                        const int sh = (int)(bit & 31u);
                        out[wi] |= (uint32_t)(v << sh);
                        if (sh + w > 32) out[wi + 1] |= (uint32_t)(v >> (32 - sh));
                        bit += (uint64_t)w;
                    }
            }
        }
    }
    if (!fill) {                                     // first call: window offsets (two slack words each)
        word_off_host[0] = 0;
        for (int64_t b = 0; b < B; ++b)
            word_off_host[b + 1] = word_off_host[b] + run_word_host[b * (int64_t)(nkeys + 1) + nkeys] + 2;
    }
    if (skipped_host) {
        skipped_host[0] = dropped;
        skipped_host[1] = odd;
    }
    return rc;
}
