// Full-ranking evaluator, exact-fp32 path: user x item score tiles fused with the counterfactual
// (TE / TIE, "rubi") epilogue, train-item masking and a per-user running top-K; then the metric
// curves.  Replaces EliMRec.predict + general_cm_fusion (reference models/EliMRec.py:96-113,
// 155-188), the Python masking loop (evaluator/backend/cpp/uni_evaluator.py:149-154) and
// cpp_evaluate_matrix / metric.h (evaluator/backend/cpp/include/evaluate.h:23-64, metric.h:17-114).
// Scores never leave the device (the reference copies a [128 x I] matrix to the host per batch).
//
// Tile: 32 users x 64 items per CTA iteration, 256 threads, (1+M) dot products of length 64 per
// (user, item) from padded shared-memory rows (conflict-free LDS.128).  Top-K: each warp owns 4 users;
// a user's current best-K list lives one entry per lane, candidates above the K-th value are
// inserted with ballot/shuffle (no shared-memory sort, no atomics).  Items are scanned in increasing
// index and insertion is strict ">" so equal scores keep the LOWEST index first (the contract's tie
// rule; the reference's partial_sort_copy order among equal keys is unspecified).
#include <float.h>
#include "common.cuh"

namespace {

constexpr int TU = 32, TI = 64, RS = 68;  // row stride (floats) in shared memory
constexpr int NTMAX = 1 + ELIMREC_MAX_MODS;
enum { RK_MEAN = 0, RK_SCORES = 1, RK_TOPK = 2 };

struct RankT {
    int U, I, n_mod, mode;
    const float* user_tab[NTMAX];
    const float* item_tab[NTMAX];
};

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

// t.mode = predict mode (0 'normal', 1 'TE', 2 'TIE') + 4 * score fusion (0 'rubi', 1 'hm', 2 'sum'); reference
// general_cm_fusion, models/EliMRec.py:171-210, eps = 1e-12 (:13).  hm / sum always use every head handed in.
__device__ __forceinline__ float fuse_fn(float x, const float* d, int n_mod, int fm) {
    if (fm == 1) {            // hm: z = ((sigmoid(x) * z_v) * z_a) * z_t ;  log(z + eps) - log1p(z)
        float z = sigm(x);
#pragma unroll
        for (int m = 0; m < ELIMREC_MAX_MODS; ++m)
            if (m < n_mod) z *= sigm(d[m + 1]);
        return logf(z + 1e-12f) - log1pf(z);
    }
    float z = x;              // sum: log(sigmoid(((x + c_v) + c_a) + c_t) + eps), raw cosines
#pragma unroll
    for (int m = 0; m < ELIMREC_MAX_MODS; ++m)
        if (m < n_mod) z += d[m + 1];
    return logf(sigm(z) + 1e-12f);
}

__device__ __forceinline__ float score_fn(const float* d, int n_mod, int mode, float mean) {
    const int pm = mode & 3, fm = mode >> 2;
    const float ui = sigm(d[0]);
    if (pm == 0) return sigm(ui);
    if (fm != 0) {
        const float te = fuse_fn(ui, d, n_mod, fm);
        if (pm == 1) return sigm(te);
        return sigm(te - fuse_fn(mean, d, n_mod, fm));
    }
    float z = ui, nd = mean;
#pragma unroll
    for (int m = 0; m < ELIMREC_MAX_MODS; ++m) {
        if (m < n_mod) {
            const float zs = sigm(d[m + 1]);
            z *= zs;      // ((ui * z_v) * z_a) * z_t   - reference evaluation order (EliMRec.py:186)
            nd *= zs;     // ((mean * z_v) * z_a) * z_t
        }
    }
    if (pm == 1) return sigm(z);
    return sigm(z - nd);
}

struct TopkArgs {
    const long long* train_ptr;
    const int* train_items;
    int K;
    int* idx;
    float* val;
};

template <int MODE>
__global__ void __launch_bounds__(256)
rank_kernel(RankT t, int n_eval, const int* __restrict__ eval_users, const float* __restrict__ mean_in,
            float* __restrict__ out, TopkArgs tk) {
    extern __shared__ float smem[];
    const int ntab = (MODE == RK_MEAN) ? 1 : 1 + ((t.mode & 3) == 0 ? 0 : t.n_mod);
    float* us = smem;                          // [NTMAX][TU][RS]
    float* is = us + NTMAX * TU * RS;          // [NTMAX][TI][RS]
    float* sc = is + NTMAX * TI * RS;          // [TU][TI+1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int u0 = blockIdx.x * TU;
    const unsigned full = 0xffffffffu;

    // stage the user rows of every table (zero rows past the end)
    for (int idx = tid; idx < ntab * TU * 16; idx += 256) {
        const int tb = idx / (TU * 16), r = (idx / 16) % TU, c4 = idx % 16;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u0 + r < n_eval) {
            const int u = __ldg(eval_users + u0 + r);
            v = __ldg(reinterpret_cast<const float4*>(t.user_tab[tb] + (long long)u * 64) + c4);
        }
        *reinterpret_cast<float4*>(us + (tb * TU + r) * RS + c4 * 4) = v;
    }
    float mean2[2] = {0.f, 0.f};
    if (MODE != RK_MEAN && (t.mode & 3) == 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (u0 + ty + 16 * h < n_eval) mean2[h] = __ldg(mean_in + u0 + ty + 16 * h);
    }
    float rowsum[2] = {0.f, 0.f};

    // per-warp top-K state (4 users per warp, one list entry per lane)
    float lv[4];
    int li[4];
    long long tp[4], te[4];
    if (MODE == RK_TOPK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            lv[q] = -INFINITY;
            li[q] = -1;
            const int gu = u0 + warp * 4 + q;
            tp[q] = te[q] = 0;
            if (gu < n_eval) {
                const int u = __ldg(eval_users + gu);
                tp[q] = __ldg(tk.train_ptr + u);
                te[q] = __ldg(tk.train_ptr + u + 1);
            }
        }
    }

    for (int i0 = 0; i0 < t.I; i0 += TI) {
        __syncthreads();  // previous tile fully consumed (also covers the user staging on the first pass)
        for (int idx = tid; idx < ntab * TI * 16; idx += 256) {
            const int tb = idx / (TI * 16), r = (idx / 16) % TI, c4 = idx % 16;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i0 + r < t.I) v = __ldg(reinterpret_cast<const float4*>(t.item_tab[tb] + (long long)(i0 + r) * 64) + c4);
            *reinterpret_cast<float4*>(is + (tb * TI + r) * RS + c4 * 4) = v;
        }
        __syncthreads();

        float acc[NTMAX][2][4];
#pragma unroll
        for (int tb = 0; tb < NTMAX; ++tb)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[tb][h][j] = 0.f;
#pragma unroll
        for (int tb = 0; tb < NTMAX; ++tb) {
            if (tb < ntab) {
                const float* ur0 = us + (tb * TU + ty) * RS;
                const float* ur1 = us + (tb * TU + ty + 16) * RS;
                const float* ir = is + (tb * TI + tx) * RS;
#pragma unroll 4
                for (int k4 = 0; k4 < 16; ++k4) {
                    const float4 a = *reinterpret_cast<const float4*>(ur0 + k4 * 4);
                    const float4 b = *reinterpret_cast<const float4*>(ur1 + k4 * 4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 v = *reinterpret_cast<const float4*>(ir + j * 16 * RS + k4 * 4);
                        acc[tb][0][j] = fmaf(a.x, v.x, acc[tb][0][j]);
                        acc[tb][0][j] = fmaf(a.y, v.y, acc[tb][0][j]);
                        acc[tb][0][j] = fmaf(a.z, v.z, acc[tb][0][j]);
                        acc[tb][0][j] = fmaf(a.w, v.w, acc[tb][0][j]);
                        acc[tb][1][j] = fmaf(b.x, v.x, acc[tb][1][j]);
                        acc[tb][1][j] = fmaf(b.y, v.y, acc[tb][1][j]);
                        acc[tb][1][j] = fmaf(b.z, v.z, acc[tb][1][j]);
                        acc[tb][1][j] = fmaf(b.w, v.w, acc[tb][1][j]);
                    }
                }
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ul = ty + 16 * h;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int il = tx + 16 * j;
                const bool valid = (i0 + il < t.I) && (u0 + ul < n_eval);
                if (MODE == RK_MEAN) {
                    if (valid) rowsum[h] += sigm(acc[0][h][j]);
                } else {
                    float d[NTMAX];
#pragma unroll
                    for (int tb = 0; tb < NTMAX; ++tb) d[tb] = acc[tb][h][j];
                    const float s = score_fn(d, t.n_mod, t.mode, mean2[h]);
                    if (MODE == RK_SCORES) {
                        if (valid) out[(long long)(u0 + ul) * t.I + i0 + il] = s;
                    } else {
                        sc[ul * (TI + 1) + il] = valid ? s : -INFINITY;
                    }
                }
            }
        }
        if (MODE == RK_TOPK) {
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int ul = warp * 4 + q;
                if (u0 + ul >= n_eval) continue;  // warp-uniform
                // training items -> -inf (uni_evaluator.py:149-154); the CSR row is sorted
                while (true) {
                    const long long e = tp[q] + lane;
                    const int it = (e < te[q]) ? __ldg(tk.train_items + e) : INT_MAX;
                    const bool in = it < i0 + TI;
                    if (in) sc[ul * (TI + 1) + (it - i0)] = -INFINITY;
                    const int n = __popc(__ballot_sync(full, in));
                    tp[q] += n;
                    if (n < 32) break;
                }
                __syncwarp();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const float v = sc[ul * (TI + 1) + half * 32 + lane];
                    float thr = __shfl_sync(full, lv[q], tk.K - 1);
                    unsigned m = __ballot_sync(full, v > thr);
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        const float cv = __shfl_sync(full, v, b);
                        if (cv > thr) {
                            const int pos = __popc(__ballot_sync(full, lane < tk.K && lv[q] >= cv));
                            const float pv = __shfl_up_sync(full, lv[q], 1);
                            const int pi = __shfl_up_sync(full, li[q], 1);
                            if (lane == pos) { lv[q] = cv; li[q] = i0 + half * 32 + b; }
                            else if (lane > pos) { lv[q] = pv; li[q] = pi; }
                            thr = __shfl_sync(full, lv[q], tk.K - 1);
                        }
                    }
                }
            }
        }
    }

    if (MODE == RK_MEAN) {
        // sum the 16 item-lanes of each user (fixed shuffle tree => deterministic), then / I
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float s = rowsum[h];
            s += __shfl_xor_sync(full, s, 8);
            s += __shfl_xor_sync(full, s, 4);
            s += __shfl_xor_sync(full, s, 2);
            s += __shfl_xor_sync(full, s, 1);
            if (tx == 0 && u0 + ty + 16 * h < n_eval) out[u0 + ty + 16 * h] = s / (float)t.I;
        }
    }
    if (MODE == RK_TOPK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int gu = u0 + warp * 4 + q;
            if (gu < n_eval && lane < tk.K) {
                tk.idx[(long long)gu * tk.K + lane] = li[q];
                tk.val[(long long)gu * tk.K + lane] = lv[q];
            }
        }
    }
}

__global__ void row_normalize_kernel(long long n_rows, const float* __restrict__ src, float* __restrict__ dst) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    const float2 v = __ldg(reinterpret_cast<const float2*>(src + r * 64) + lane);
    const float c = fmaxf(sqrtf(warp_sum(v.x * v.x + v.y * v.y)), 1e-12f);
    reinterpret_cast<float2*>(dst + r * 64)[lane] = make_float2(v.x / c, v.y / c);
}

// top-K of an explicit [rows x cols] matrix, one warp per row (arg_top_k_2d replacement)
__global__ void topk_matrix_kernel(int n_rows, int n_cols, const float* __restrict__ scores, int K,
                                   int* __restrict__ idx, float* __restrict__ val) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    if (r >= n_rows) return;
    float lv = -INFINITY;
    int li = -1;
    const float* row = scores + (long long)r * n_cols;
    for (int c0 = 0; c0 < n_cols; c0 += 32) {
        const float v = (c0 + lane < n_cols) ? __ldg(row + c0 + lane) : -INFINITY;
        float thr = __shfl_sync(full, lv, K - 1);
        unsigned m = __ballot_sync(full, v > thr);
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const float cv = __shfl_sync(full, v, b);
            if (cv > thr) {
                const int pos = __popc(__ballot_sync(full, lane < K && lv >= cv));
                const float pv = __shfl_up_sync(full, lv, 1);
                const int pi = __shfl_up_sync(full, li, 1);
                if (lane == pos) { lv = cv; li = c0 + b; }
                else if (lane > pos) { lv = pv; li = pi; }
                thr = __shfl_sync(full, lv, K - 1);
            }
        }
    }
    if (lane < K) {
        idx[(long long)r * K + lane] = li;
        val[(long long)r * K + lane] = lv;
    }
}

// The same for 32 < K <= 128: NL = ceil(K / 32) list entries per lane, lane l holds the sorted positions [l NL, (l+1) NL).
// Strict `>` against the current K-th value in increasing column order => ties keep the lowest index, like the K <= 32 path.
template <int NL>
__global__ void topk_matrix_big_kernel(int n_rows, int n_cols, const float* __restrict__ scores, int K,
                                       int* __restrict__ idx, float* __restrict__ val) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    if (r >= n_rows) return;
    float lv[NL];
    int li[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) { lv[j] = -INFINITY; li[j] = -1; }
    const int kl = (K - 1) / NL, kj = (K - 1) % NL;      // where the K-th entry lives
    const float* row = scores + (long long)r * n_cols;
    auto kth = [&]() {
        float t = lv[0];
#pragma unroll
        for (int j = 1; j < NL; ++j) if (j == kj) t = lv[j];
        return __shfl_sync(full, t, kl);
    };
    for (int c0 = 0; c0 < n_cols; c0 += 32) {
        const float v = (c0 + lane < n_cols) ? __ldg(row + c0 + lane) : -INFINITY;
        float thr = kth();
        unsigned m = __ballot_sync(full, v > thr);
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const float cv = __shfl_sync(full, v, b);
            if (cv > thr) {
                int cnt = 0;                              // entries (among the first K) that stay in front of the candidate
#pragma unroll
                for (int j = 0; j < NL; ++j) cnt += (lane * NL + j < K && lv[j] >= cv) ? 1 : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(full, cnt, o);
                const int pos = cnt;
                const float pv = __shfl_up_sync(full, lv[NL - 1], 1);
                const int pi = __shfl_up_sync(full, li[NL - 1], 1);
#pragma unroll
                for (int j = NL - 1; j >= 0; --j) {
                    const int p = lane * NL + j;
                    if (p > pos) {
                        lv[j] = (j > 0) ? lv[j - 1] : pv;
                        li[j] = (j > 0) ? li[j - 1] : pi;
                    } else if (p == pos) {
                        lv[j] = cv;
                        li[j] = c0 + b;
                    }
                }
                thr = kth();
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        const int p = lane * NL + j;
        if (p < K) {
            idx[(long long)r * K + p] = li[j];
            val[(long long)r * K + p] = lv[j];
        }
    }
}

// scores[r, train items of users[r]] = -inf   (uni_evaluator.py:149-154), for the explicit-matrix path
__global__ void mask_train_kernel(int n_rows, int n_cols, const int* __restrict__ users, const long long* __restrict__ tptr,
                                  const int* __restrict__ titems, float* __restrict__ scores) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    const int u = __ldg(users + r);
    for (long long e = __ldg(tptr + u) + lane; e < __ldg(tptr + u + 1); e += 32) {
        const int it = __ldg(titems + e);
        if (it >= 0 && it < n_cols) scores[(long long)r * n_cols + it] = -INFINITY;
    }
}

constexpr int METRIC_MAX_K = 128;

struct MetricArgs {
    int n_metrics;
    int ids[5];
    double inv_log2[METRIC_MAX_K];
};

// metric.h:17-114 for 32 < K <= 128: same accumulator types and order, hit flags in ceil(K / 32) ballots, curves by lane 0
__global__ void metric_rows_big_kernel(int n_eval, int K, const int* __restrict__ topk, const long long* __restrict__ tptr,
                                       const int* __restrict__ titems, MetricArgs ma, float* __restrict__ rows) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    if (r >= n_eval) return;
    const long long b = __ldg(tptr + r), e = __ldg(tptr + r + 1);
    const int tl = (int)(e - b);
    unsigned hm[METRIC_MAX_K / 32];
#pragma unroll
    for (int c = 0; c < METRIC_MAX_K / 32; ++c) {
        bool hit = false;
        const int i = c * 32 + lane;
        if (i < K) {
            const int id = __ldg(topk + (long long)r * K + i);
            long long lo = b, hi = e;
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if (__ldg(titems + mid) < id) lo = mid + 1; else hi = mid;
            }
            hit = (lo < e) && (__ldg(titems + lo) == id);
        }
        hm[c] = __ballot_sync(full, hit);
    }
    if (lane != 0) return;
    float* out = rows + (long long)r * ma.n_metrics * K;
    for (int j = 0; j < ma.n_metrics; ++j) {
        const int id = ma.ids[j];
        float* o = out + j * K;
        int hits = 0, first = K;
        float dcg = 0.f, idcg = 0.f, sum_pre = 0.f;
        for (int i = 0; i < K; ++i) {
            const bool h = (hm[i >> 5] >> (i & 31)) & 1u;
            if (h) {
                hits += 1;
                if (first == K) first = i;
            }
            if (id == 1) o[i] = (float)(1.0 * hits / (i + 1));
            else if (id == 2) o[i] = (float)(1.0 * hits / (double)tl);
            else if (id == 4) {
                if (h) dcg = (float)((double)dcg + ma.inv_log2[i]);
                if (i < tl) idcg = (float)((double)idcg + ma.inv_log2[i]);
                o[i] = dcg / idcg;
            } else if (id == 3) {
                if (h) sum_pre += (float)(1.0 * hits / (i + 1));
                o[i] = (hits == 0) ? 0.f : sum_pre / (float)hits;
            } else {
                o[i] = (i >= first) ? (float)(1.0 / (first + 1)) : 0.f;
            }
        }
    }
}

// metric.h:17-114 with the reference's accumulator types (int hits, float DCG/iDCG/sum_pre, double
// addends, float stores).  One warp per evaluated user; lane i owns rank i.
__global__ void metric_rows_kernel(int n_eval, int K, const int* __restrict__ topk, const long long* __restrict__ tptr,
                                   const int* __restrict__ titems, MetricArgs ma, float* __restrict__ rows) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    if (r >= n_eval) return;
    const long long b = __ldg(tptr + r), e = __ldg(tptr + r + 1);
    const int tl = (int)(e - b);
    bool hit = false;
    if (lane < K) {
        const int id = __ldg(topk + (long long)r * K + lane);
        long long lo = b, hi = e;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (__ldg(titems + mid) < id) lo = mid + 1; else hi = mid;
        }
        hit = (lo < e) && (__ldg(titems + lo) == id);
    }
    const unsigned hm = __ballot_sync(full, hit);
    const int hits = __popc(hm & (lane == 31 ? 0xffffffffu : ((2u << lane) - 1u)));
    float* out = rows + (long long)r * ma.n_metrics * K;
    for (int j = 0; j < ma.n_metrics; ++j) {
        const int id = ma.ids[j];
        float* o = out + j * K;
        if (id == 1) {
            if (lane < K) o[lane] = (float)(1.0 * hits / (lane + 1));
        } else if (id == 2) {
            if (lane < K) o[lane] = (float)(1.0 * hits / (double)tl);
        } else if (lane == 0) {
            if (id == 4) {
                float dcg = 0.f, idcg = 0.f;
                for (int i = 0; i < K; ++i) {
                    if ((hm >> i) & 1u) dcg = (float)((double)dcg + ma.inv_log2[i]);
                    if (i < tl) idcg = (float)((double)idcg + ma.inv_log2[i]);
                    o[i] = dcg / idcg;
                }
            } else if (id == 3) {
                int h = 0;
                float sum_pre = 0.f;
                for (int i = 0; i < K; ++i) {
                    if ((hm >> i) & 1u) {
                        h += 1;
                        const float pre = (float)(1.0 * h / (i + 1));
                        sum_pre += pre;
                    }
                    o[i] = (h == 0) ? 0.f : sum_pre / (float)h;
                }
            } else {  // 5: mrr
                const int f = hm ? (__ffs(hm) - 1) : K;
                for (int i = 0; i < K; ++i) o[i] = (i >= f) ? (float)(1.0 / (f + 1)) : 0.f;
            }
        }
    }
}

// sums[c] += sum_r rows[r][c] in double; one block per column, fixed tree => deterministic
__global__ void metric_colsum_kernel(int n_rows, int n_cols, const float* __restrict__ rows, double* __restrict__ sums) {
    __shared__ double red[256];
    const int c = blockIdx.x;
    double s = 0.0;
    for (int r = threadIdx.x; r < n_rows; r += 256) s += (double)rows[(long long)r * n_cols + c];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[c] += red[0];
}

int fill_tables(const elimrec_rank_tables_t* t, RankT* r) {
    if (t == nullptr) return -1;
    if (t->n_mod < 0 || t->n_mod > ELIMREC_MAX_MODS || t->mode < 0 || (t->mode & 3) > 2 || (t->mode >> 2) > 2) return -1;
    r->U = t->num_users; r->I = t->num_items; r->n_mod = t->n_mod; r->mode = t->mode;
    for (int i = 0; i < NTMAX; ++i) { r->user_tab[i] = nullptr; r->item_tab[i] = nullptr; }
    r->user_tab[0] = t->f_user;
    r->item_tab[0] = t->f_item;
    for (int m = 0; m < t->n_mod; ++m) {
        r->user_tab[m + 1] = t->s_user[m];
        r->item_tab[m + 1] = t->s_item[m];
    }
    return 0;
}

constexpr size_t RANK_SMEM = (size_t)(NTMAX * TU * RS + NTMAX * TI * RS + TU * (TI + 1)) * sizeof(float);

template <int MODE>
int launch_rank(const RankT& r, int n_eval, const int* eval_users, const float* mean_in, float* out, TopkArgs tk,
                cudaStream_t st) {
    static bool configured = false;  // benign race: idempotent attribute set
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(rank_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RANK_SMEM);
        if (e != cudaSuccess) { elimrec_set_error("rank: cannot opt in to %zu B shared memory: %s", RANK_SMEM, cudaGetErrorString(e)); return -3; }
        configured = true;
    }
    const int blocks = (n_eval + TU - 1) / TU;
    rank_kernel<MODE><<<blocks, 256, RANK_SMEM, st>>>(r, n_eval, eval_users, mean_in, out, tk);
    return 0;
}

}  // namespace

ELIMREC_API int elimrec_row_normalize(int64_t n_rows, const float* src, float* dst, elimrec_stream_t stream) {
    if (n_rows <= 0) return 0;
    row_normalize_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, er_stream(stream)>>>(n_rows, src, dst);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_rank_rowmean(const elimrec_rank_tables_t* t, int n_eval, const int32_t* eval_users,
                                     float* ui_mean, elimrec_stream_t stream) {
    RankT r;
    ER_CHECK_ARG(fill_tables(t, &r) == 0, "bad table descriptor");
    if (n_eval <= 0) return 0;
    TopkArgs tk{};
    if (launch_rank<RK_MEAN>(r, n_eval, eval_users, nullptr, ui_mean, tk, er_stream(stream)) != 0) return -3;
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_rank_scores(const elimrec_rank_tables_t* t, int n_eval, const int32_t* eval_users,
                                    const float* ui_mean, float* scores, elimrec_stream_t stream) {
    RankT r;
    ER_CHECK_ARG(fill_tables(t, &r) == 0, "bad table descriptor");
    ER_CHECK_ARG((t->mode & 3) != 2 || ui_mean != nullptr, "TIE needs ui_mean");
    if (n_eval <= 0) return 0;
    TopkArgs tk{};
    if (launch_rank<RK_SCORES>(r, n_eval, eval_users, ui_mean, scores, tk, er_stream(stream)) != 0) return -3;
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_rank_topk(const elimrec_rank_tables_t* t, int n_eval, const int32_t* eval_users,
                                  const float* ui_mean, const int64_t* train_ptr, const int32_t* train_items, int K,
                                  int32_t* topk_idx, float* topk_val, elimrec_stream_t stream) {
    RankT r;
    ER_CHECK_ARG(fill_tables(t, &r) == 0, "bad table descriptor");
    ER_CHECK_ARG(K >= 1 && K <= 32, "K must be in [1, 32]");
    ER_CHECK_ARG((t->mode & 3) != 2 || ui_mean != nullptr, "TIE needs ui_mean");
    if (n_eval <= 0) return 0;
    TopkArgs tk{(const long long*)train_ptr, train_items, K, topk_idx, topk_val};
    if (launch_rank<RK_TOPK>(r, n_eval, eval_users, ui_mean, nullptr, tk, er_stream(stream)) != 0) return -3;
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_topk_matrix(int n_rows, int n_cols, const float* scores, int K, int32_t* topk_idx,
                                    float* topk_val, elimrec_stream_t stream) {
    ER_CHECK_ARG(K >= 1 && K <= METRIC_MAX_K, "K must be in [1, 128]");
    if (n_rows <= 0) return 0;
    cudaStream_t st = er_stream(stream);
    const int blocks = (n_rows + 7) / 8;
    if (K <= 32) topk_matrix_kernel<<<blocks, 256, 0, st>>>(n_rows, n_cols, scores, K, topk_idx, topk_val);
    else if (K <= 64) topk_matrix_big_kernel<2><<<blocks, 256, 0, st>>>(n_rows, n_cols, scores, K, topk_idx, topk_val);
    else topk_matrix_big_kernel<4><<<blocks, 256, 0, st>>>(n_rows, n_cols, scores, K, topk_idx, topk_val);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_mask_train(int n_rows, int n_cols, const int32_t* users, const int64_t* train_ptr,
                                   const int32_t* train_items, float* scores, elimrec_stream_t stream) {
    ER_CHECK_ARG(users != nullptr && train_ptr != nullptr && scores != nullptr, "NULL buffer");
    if (n_rows <= 0) return 0;
    mask_train_kernel<<<(n_rows + 7) / 8, 256, 0, er_stream(stream)>>>(n_rows, n_cols, users, (const long long*)train_ptr, train_items,
                                                                      scores);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_metric_rows(int n_eval, int K, const int32_t* topk_idx, const int64_t* truth_ptr,
                                    const int32_t* truth_items, int n_metrics, const int32_t* metric_ids_host,
                                    const double* inv_log2_host, float* rows, double* sums, elimrec_stream_t stream) {
    ER_CHECK_ARG(K >= 1 && K <= METRIC_MAX_K, "K must be in [1, 128]");
    ER_CHECK_ARG(n_metrics >= 1 && n_metrics <= 5, "n_metrics must be in [1, 5]");
    if (n_eval <= 0) return 0;
    MetricArgs ma{};
    ma.n_metrics = n_metrics;
    for (int j = 0; j < n_metrics; ++j) {
        ER_CHECK_ARG(metric_ids_host[j] >= 1 && metric_ids_host[j] <= 5, "unknown metric id");
        ma.ids[j] = metric_ids_host[j];
    }
    for (int i = 0; i < K; ++i) ma.inv_log2[i] = inv_log2_host[i];
    cudaStream_t st = er_stream(stream);
    if (K <= 32)
        metric_rows_kernel<<<(n_eval + 7) / 8, 256, 0, st>>>(n_eval, K, topk_idx, (const long long*)truth_ptr, truth_items, ma, rows);
    else
        metric_rows_big_kernel<<<(n_eval + 7) / 8, 256, 0, st>>>(n_eval, K, topk_idx, (const long long*)truth_ptr, truth_items, ma,
                                                                 rows);
    ER_LAUNCH_CHECK();
    if (sums != nullptr) {
        metric_colsum_kernel<<<n_metrics * K, 256, 0, st>>>(n_eval, n_metrics * K, rows, sums);
        ER_LAUNCH_CHECK();
    }
    return 0;
}
