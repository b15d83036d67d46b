// BPR triple sampling.
//  * compat (host): bit-exact replay of the reference epoch sampler (data/sampler.py:93-126,
//    util/cython/random_choice.pyx:12-62).  The reference consumes libc rand(), never seeded, so the
//    stream is glibc's TYPE_3 additive-feedback generator from seed 1; it is re-implemented here
//    (re-entrant, caller-owned state) instead of touching libc's global state.  The algorithm is
//    sequential by construction (data-dependent rejection counts, first-occurrence user order).
//  * device: one thread per triple, counter-based Philox4x32-10 - same distribution (user uniform over
//    users with train items, positive uniform over the user's items, negative uniform over non-train
//    items by rejection), embarrassingly parallel.  CPU restatement: oracle/philox_sampler.py.
#include <stdlib.h>
#include <vector>
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// glibc random_r() TYPE_3 (x^31 + x^3 + 1), as used by rand(): state words [0..30], [31]=front, [32]=rear
// ---------------------------------------------------------------------------------------------
ELIMREC_API uint32_t elimrec_compat_rng_next(uint32_t* st) {
    uint32_t f = st[31], r = st[32];
    st[f] += st[r];
    const uint32_t out = st[f] >> 1;
    st[31] = (f + 1 == 31) ? 0 : f + 1;
    st[32] = (r + 1 == 31) ? 0 : r + 1;
    return out;
}

ELIMREC_API void elimrec_compat_rng_seed(uint32_t* st, uint32_t seed) {
    if (seed == 0) seed = 1;
    int32_t word = (int32_t)seed;
    st[0] = (uint32_t)word;
    for (int i = 1; i < 31; ++i) {
        const long hi = word / 127773, lo = word % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        word = (int32_t)w;
        st[i] = (uint32_t)word;
    }
    st[31] = 3;  // front = rear + SEP_3
    st[32] = 0;
    for (int i = 0; i < 310; ++i) (void)elimrec_compat_rng_next(st);
}

static inline unsigned long long compat_llrand(uint32_t* st) {  // random_choice.pyx:12-17
    unsigned long long r = 0;
    for (int i = 0; i < 5; ++i) r = (r << 15) | (unsigned long long)(elimrec_compat_rng_next(st) & 0x7FFF);
    return r;
}

static inline bool sorted_contains(const int32_t* a, int n, int x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo < n && a[lo] == x;
}

ELIMREC_API int elimrec_sample_epoch_compat(uint32_t* st, int32_t n_tu, const int32_t* user_ids,
                                            const int64_t* row_ptr, const int32_t* items, int32_t num_items,
                                            int64_t num_samples, int64_t* out_u, int64_t* out_p, int64_t* out_n) {
    ER_CHECK_ARG(n_tu > 0 && num_items > 0 && num_samples >= 0, "empty input");
    try {
        std::vector<int32_t> slot(num_samples), order;
        std::vector<int64_t> cnt(n_tu, 0), off(n_tu, 0);
        order.reserve(n_tu);
        for (int64_t k = 0; k < num_samples; ++k) {  // sampler.py:100-107
            const int32_t s = (int32_t)(compat_llrand(st) % (unsigned long long)n_tu);
            slot[k] = s;
            if (cnt[s]++ == 0) order.push_back(s);
        }
        int64_t acc = 0;
        for (int32_t s : order) { off[s] = acc; acc += cnt[s]; }
        std::vector<int32_t> posd(num_samples), negd(num_samples);
        for (int32_t s : order) {  // sampler.py:111-119, users in first-occurrence order
            const int32_t* it = items + row_ptr[s];
            const int deg = (int)(row_ptr[s + 1] - row_ptr[s]);
            if (num_items <= deg) {
                elimrec_set_error("elimrec_sample_epoch_compat: user %d owns every item", user_ids[s]);
                return -1;
            }
            for (int64_t c = 0; c < cnt[s]; ++c) posd[off[s] + c] = it[compat_llrand(st) % (unsigned long long)deg];
            for (int64_t got = 0; got < cnt[s];) {
                const int a = (int)(compat_llrand(st) % (unsigned long long)num_items);
                if (!sorted_contains(it, deg, a)) negd[off[s] + got++] = a;
            }
        }
        for (int64_t k = 0; k < num_samples; ++k) {  // sampler.py:123-124: list.pop() takes from the END
            const int32_t s = slot[k];
            const int64_t j = --cnt[s];
            out_u[k] = user_ids[s];
            out_p[k] = posd[off[s] + j];
            out_n[k] = negd[off[s] + j];
        }
    } catch (...) {
        elimrec_set_error("elimrec_sample_epoch_compat: out of memory");
        return -4;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// batch_index (may be NULL): device-resident batch counter; triple `t` of this launch is sample number
// (*batch_index) * n + t of the epoch's stream, so a CUDA graph that contains this launch draws a NEW batch at every replay
// (the counter is the optimizer's step counter, advanced by elimrec_adam_tick at the end of each step).
__global__ void sample_triples_kernel(unsigned long long seed, unsigned long long epoch, const long long* __restrict__ batch_index,
                                      long long n, int n_tu,
                                      const int* __restrict__ user_ids, const long long* __restrict__ row_ptr,
                                      const int* __restrict__ items, int num_items, long long* __restrict__ ou,
                                      long long* __restrict__ op, long long* __restrict__ on) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long k = (batch_index != nullptr ? *batch_index * n : 0) + t;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t c0 = (uint32_t)k, c1 = (uint32_t)((unsigned long long)k >> 32), c3 = (uint32_t)epoch;
    uint32_t o[4];
    philox4x32_10(c0, c1, 0u, c3, k0, k1, o);
    const int s = (int)((((unsigned long long)o[0] << 32) | o[1]) % (unsigned long long)n_tu);
    const long long b = row_ptr[s];
    const int deg = (int)(row_ptr[s + 1] - b);
    const int* it = items + b;
    const int p = it[(((unsigned long long)o[2] << 32) | o[3]) % (unsigned long long)deg];
    int ng = -1;
    for (uint32_t j = 1; ng < 0; ++j) {
        philox4x32_10(c0, c1, j, c3, k0, k1, o);
#pragma unroll
        for (int h = 0; h < 2 && ng < 0; ++h) {
            const int a = (int)((((unsigned long long)o[2 * h] << 32) | o[2 * h + 1]) % (unsigned long long)num_items);
            int lo = 0, hi = deg;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (it[mid] < a) lo = mid + 1; else hi = mid;
            }
            if (!(lo < deg && it[lo] == a)) ng = a;
        }
    }
    ou[t] = user_ids[s];
    op[t] = p;
    on[t] = ng;
}

}  // namespace

ELIMREC_API int elimrec_sample_triples_device(uint64_t seed, uint64_t epoch, int64_t num_samples, int32_t n_tu,
                                              const int32_t* user_ids, const int64_t* row_ptr, const int32_t* items,
                                              int32_t num_items, int64_t* out_users, int64_t* out_pos, int64_t* out_neg,
                                              elimrec_stream_t stream) {
    ER_CHECK_ARG(n_tu > 0 && num_items > 0, "empty input");
    if (num_samples <= 0) return 0;
    sample_triples_kernel<<<(unsigned)((num_samples + 255) / 256), 256, 0, er_stream(stream)>>>(
        seed, epoch, nullptr, num_samples, n_tu, user_ids, (const long long*)row_ptr, items, num_items, (long long*)out_users,
        (long long*)out_pos, (long long*)out_neg);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_sample_batch_device(uint64_t seed, uint64_t epoch, const int64_t* batch_index_dev, int64_t batch_size,
                                            int32_t n_tu, const int32_t* user_ids, const int64_t* row_ptr, const int32_t* items,
                                            int32_t num_items, int64_t* out_users, int64_t* out_pos, int64_t* out_neg,
                                            elimrec_stream_t stream) {
    ER_CHECK_ARG(n_tu > 0 && num_items > 0 && batch_index_dev != nullptr, "empty input");
    if (batch_size <= 0) return 0;
    sample_triples_kernel<<<(unsigned)((batch_size + 255) / 256), 256, 0, er_stream(stream)>>>(
        seed, epoch, (const long long*)batch_index_dev, batch_size, n_tu, user_ids, (const long long*)row_ptr, items, num_items,
        (long long*)out_users, (long long*)out_pos, (long long*)out_neg);
    ER_LAUNCH_CHECK();
    return 0;
}
