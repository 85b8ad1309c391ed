/* TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * Exhaustive K-nearest-neighbour search used by oracle/focus_oracle.py so that the
 * 480x640 / 19200-trajectory case finishes in seconds on the host cores.  It restates
 * the reference's KeOps reduction `dist.argKmin(K, dim=2)` / `dist.Kmin(K, axis=2)`
 * (src/losses/focus.py:129-137,159 upstream): for every LUT query, the K trajectories
 * with the smallest distance, ascending, lowest trajectory index first on equal distance.
 *
 * Distance arithmetic (float32, no FMA contraction: compile with -ffp-contract=off):
 *   l2: fl(fl(dy*dy) + fl(dx*dx))      l1: fl(|dy| + |dx|)      dy = gy - py, dx = gx - px
 *
 * points [nslab, n, 2] f32 (y, x) ; grid [q, 2] f32 ; out ind [nslab, q, K] int64,
 * dist [nslab, q, K] f32.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_knn_bruteforce(const float *points, const float *grid, int64_t nslab, int64_t n,
                          int64_t q, int K, int l1, int64_t *ind, float *dist, int nthreads)
{
    if (K <= 0 || K > n) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    int64_t total = nslab * q;
#pragma omp parallel
    {
        float *bd = (float *)malloc(sizeof(float) * (size_t)K);
        int64_t *bi = (int64_t *)malloc(sizeof(int64_t) * (size_t)K);
#pragma omp for schedule(dynamic, 64)
        for (int64_t w = 0; w < total; ++w) {
            int64_t slab = w / q, c = w % q;
            const float *p = points + slab * n * 2;
            float gy = grid[2 * c], gx = grid[2 * c + 1];
            int cnt = 0;
            for (int64_t j = 0; j < n; ++j) {
                volatile float dy = gy - p[2 * j], dx = gx - p[2 * j + 1];
                float d;
                if (l1) {
                    d = fabsf(dy) + fabsf(dx);
                } else {
                    volatile float a = dy * dy, b = dx * dx;
                    d = a + b;
                }
                if (cnt == K && !(d < bd[K - 1])) continue;     /* ties keep the earlier index */
                int pos = cnt < K ? cnt : K - 1;
                while (pos > 0 && d < bd[pos - 1]) {
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = d;
                bi[pos] = j;
                if (cnt < K) ++cnt;
            }
            for (int k = 0; k < K; ++k) {
                ind[w * K + k] = bi[k];
                dist[w * K + k] = bd[k];
            }
        }
        free(bd);
        free(bi);
    }
    return 0;
}
