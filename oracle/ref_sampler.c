/* ORACLE (test infrastructure).  Plain-C restatement of the reference's epoch sampler, driven by
 * the SAME libc rand() stream the reference consumes:
 *   util/cython/random_choice.pyx:12-17  llrand(): five rand()&0x7FFF packed into a wrapping u64
 *   util/cython/random_choice.pyx:20-62  randint_choice(): `llrand() % high`, rejection of excluded ids
 *   data/sampler.py:93-126               _pairwise_sampling_v2(): user draws, first-occurrence order,
 *                                        per-user pos then neg draws, pop() from the END
 * Pinned by tests/golden (epoch_users / epoch_pos / epoch_neg recorded from the reference itself).
 */
#include <stdlib.h>
#include <string.h>

static unsigned long long llrand(void) {
    unsigned long long r = 0;
    for (int i = 0; i < 5; ++i) r = (r << 15) | (unsigned long long)(rand() & 0x7FFF);
    return r;
}

static int contains(const int* a, int n, int x) { /* sorted ascending */
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
    return lo < n && a[lo] == x;
}

void ref_srand(unsigned s) { srand(s); }

/* user_ids[n_tu]: users that have train items, in dict order; row_ptr[n_tu+1]/items: their sorted
 * train items.  Outputs: num_samples triples.  Returns 0, or -1 if a user owns every item. */
int ref_sample_epoch(int n_tu, const int* user_ids, const long long* row_ptr, const int* items, int num_items,
                     long long num_samples, int* out_u, int* out_p, int* out_n) {
    int* slot = (int*)malloc(sizeof(int) * num_samples);          /* index into user_ids per sample */
    long long* cnt = (long long*)calloc(n_tu, sizeof(long long));
    int* order = (int*)malloc(sizeof(int) * n_tu);
    int n_order = 0;
    for (long long k = 0; k < num_samples; ++k) {                   /* sampler.py:100-107 */
        int s = (int)(llrand() % (unsigned long long)n_tu);
        slot[k] = s;
        if (cnt[s]++ == 0) order[n_order++] = s;
    }
    long long* off = (long long*)malloc(sizeof(long long) * n_tu);
    long long acc = 0;
    for (int j = 0; j < n_order; ++j) { off[order[j]] = acc; acc += cnt[order[j]]; }
    int* posd = (int*)malloc(sizeof(int) * num_samples);
    int* negd = (int*)malloc(sizeof(int) * num_samples);
    for (int j = 0; j < n_order; ++j) {                             /* sampler.py:111-119 */
        int s = order[j];
        const int* it = items + row_ptr[s];
        int deg = (int)(row_ptr[s + 1] - row_ptr[s]);
        if (num_items <= deg) return -1;
        for (long long c = 0; c < cnt[s]; ++c) posd[off[s] + c] = it[llrand() % (unsigned long long)deg];
        long long got = 0;
        while (got < cnt[s]) {
            int a = (int)(llrand() % (unsigned long long)num_items);
            if (!contains(it, deg, a)) negd[off[s] + got++] = a;
        }
    }
    for (long long k = 0; k < num_samples; ++k) {                   /* sampler.py:123-124: pop() */
        int s = slot[k];
        long long j = --cnt[s];
        out_u[k] = user_ids[s];
        out_p[k] = posd[off[s] + j];
        out_n[k] = negd[off[s] + j];
    }
    free(slot); free(cnt); free(order); free(off); free(posd); free(negd);
    return 0;
}
