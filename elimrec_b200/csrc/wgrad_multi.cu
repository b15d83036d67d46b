// Every weight / bias gradient of a training step whose contraction runs over the 3B instance rows, in ONE launch + one
// fixed-order reduction: the fusion Linears (embedding_{user,item}_after_GCN), the single-modal heads (s_dense_*) and - in the
// linear schedule - the packed modality projections d[W_m | b_m] = dO_m[inst]^T Zbar_m[inst]
// (reference: the autograd of models/EliMRec.py:233-236, 146-151, 261-270 restricted to the rows where it is non-zero).
//
//   out_p [64 x K_p] = g * A_p[r0:r1, 0:64]^T  B_p[r0:r1, 0:K_p]        bias_p [64] = g * column sums of A_p[r0:r1]
//
// These are "skinny" reductions (1.2 GFLOP at the Tiktok shape, 6 K rows) that used to take ten launches (chunked outer
// products + reductions, hi/lo operand splits + three tensor-core weight-gradient launches + their reductions) and ~200 us
// of kernel time beside the propagation backward.  Exact fp32 FFMA, deterministic: a CTA owns a 64 x 64 output tile and one
// of `splits` row ranges (register-staged double buffering, 64 x 128 tile, 8 x 8 micro-tile), partial tiles go to scratch and are
// summed in split order.
#include "common.cuh"

namespace {

constexpr int WG_BR = 16;          // rows per staged chunk
constexpr int WG_TN = 128;         // output columns per tile (B columns); 64 output rows (A columns)

struct WgProblem {
    const float* A;
    long long lda;
    const float* B;
    long long ldb;
    int K;
    int r0, r1;
    float* out;
    long long ldo;
    float* bias;        // NULL: no bias gradient
    int scale_by_g;
    int tile0;          // first tile of this problem
};

struct WgArgs {
    int n_prob;
    WgProblem p[ELIMREC_WGRAD_MAX_PROBLEMS];
    int n_tiles;
    int splits;
    float* ws;          // [n_tiles][splits][64 x 128] partial tiles, then [n_tiles][splits][64] partial column sums
    const float* gscale;
};

__device__ __forceinline__ int wg_find(const WgArgs& a, int tile) {
    int q = 0;
    while (q + 1 < a.n_prob && tile >= a.p[q + 1].tile0) ++q;
    return q;
}

// 128 threads, 8 x 8 outputs per thread: thread (ty, tx) owns output rows {4 ty .. 4 ty + 3, 32 + 4 ty ..} (A columns) and output
// columns {4 tx .. 4 tx + 3, 64 + 4 tx ..} (B columns) - two conflict-free LDS.128 per operand and 64 FMAs per staged row, so the loop is
// bound by the FP32 pipe and not by shared-memory bandwidth (a 4 x 4 micro-tile spends as many cycles loading as multiplying).
__global__ void __launch_bounds__(128) wgrad_multi_kernel(const __grid_constant__ WgArgs a) {
    __shared__ __align__(16) float As[2][WG_BR][64];
    __shared__ __align__(16) float Bs[2][WG_BR][WG_TN];
    const int tile = blockIdx.x, split = blockIdx.y;
    const WgProblem& pr = a.p[wg_find(a, tile)];
    const int c0 = (tile - pr.tile0) * WG_TN;              // first column of B / out handled here
    const int rows = pr.r1 - pr.r0;
    int chunk = (rows + a.splits - 1) / a.splits;
    chunk = (chunk + WG_BR - 1) / WG_BR * WG_BR;
    const int rb = pr.r0 + split * chunk, re = min(pr.r1, rb + chunk);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;
    const bool want_bias = pr.bias != nullptr && c0 == 0;
    const bool b_vec = (pr.ldb % 4 == 0) && ((reinterpret_cast<unsigned long long>(pr.B) & 15) == 0);

    // staging: A chunk = 16 x 64 floats = 256 float4 (2 per thread), B chunk = 16 x 128 = 512 float4 (4 per thread)
    float4 ra[2], rbv[4];
    auto load = [&](int r) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int f = tid + 128 * q, row = r + (f >> 4), col = (f & 15) * 4;
            ra[q] = (row < re) ? __ldg(reinterpret_cast<const float4*>(pr.A + (long long)row * pr.lda + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int f = tid + 128 * q, row = r + (f >> 5), col = c0 + (f & 31) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < re && col < pr.K) {
                const float* bp = pr.B + (long long)row * pr.ldb + col;
                if (b_vec && col + 4 <= pr.K) {
                    v = __ldg(reinterpret_cast<const float4*>(bp));
                } else {
                    v.x = __ldg(bp);
                    if (col + 1 < pr.K) v.y = __ldg(bp + 1);
                    if (col + 2 < pr.K) v.z = __ldg(bp + 2);
                    if (col + 3 < pr.K) v.w = __ldg(bp + 3);
                }
            }
            rbv[q] = v;
        }
    };
    auto store = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int f = tid + 128 * q;
            *reinterpret_cast<float4*>(&As[buf][f >> 4][(f & 15) * 4]) = ra[q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int f = tid + 128 * q;
            *reinterpret_cast<float4*>(&Bs[buf][f >> 5][(f & 31) * 4]) = rbv[q];
        }
    };

    if (rb < re) {
        load(rb);
        store(0);
        __syncthreads();
        int buf = 0;
        for (int r = rb; r < re; r += WG_BR) {
            const bool more = r + WG_BR < re;
            if (more) load(r + WG_BR);
#pragma unroll
            for (int kk = 0; kk < WG_BR; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][32 + ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            if (want_bias && tid < 64) {
#pragma unroll
                for (int kk = 0; kk < WG_BR; ++kk) bsum += As[buf][kk][tid];
            }
            if (more) {
                store(buf ^ 1);        // the other buffer: nobody reads it during this iteration
                __syncthreads();
                buf ^= 1;
            }
        }
    }
    // partial tile, row-major [64][128]
    float* part = a.ws + ((long long)tile * a.splits + split) * (64 * WG_TN);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int o = (i < 4) ? ty * 4 + i : 32 + ty * 4 + (i - 4);
        *reinterpret_cast<float4*>(part + o * WG_TN + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(part + o * WG_TN + 64 + tx * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
    if (want_bias && tid < 64)
        a.ws[(long long)a.n_tiles * a.splits * (64 * WG_TN) + ((long long)tile * a.splits + split) * 64 + tid] = bsum;
}

// grid (tiles, 32): CTA y sums 256 elements of the tile over the splits, in split order
__global__ void __launch_bounds__(256) wgrad_multi_reduce_kernel(const __grid_constant__ WgArgs a) {
    constexpr int TE = 64 * WG_TN;
    const int tile = blockIdx.x;
    const WgProblem& pr = a.p[wg_find(a, tile)];
    const int c0 = (tile - pr.tile0) * WG_TN;
    const float g = (pr.scale_by_g && a.gscale != nullptr) ? __ldg(a.gscale) : 1.f;
    const float* part = a.ws + (long long)tile * a.splits * TE;
    {
        const int e = blockIdx.y * 256 + threadIdx.x;
        const int o = e / WG_TN, c = e % WG_TN;
        if (c0 + c < pr.K) {
            float s = 0.f;
            int z = 0;
            for (; z + 4 <= a.splits; z += 4) {
                const float t0 = part[(long long)z * TE + e], t1 = part[(long long)(z + 1) * TE + e];
                const float t2 = part[(long long)(z + 2) * TE + e], t3 = part[(long long)(z + 3) * TE + e];
                s = (((s + t0) + t1) + t2) + t3;
            }
            for (; z < a.splits; ++z) s += part[(long long)z * TE + e];
            pr.out[(long long)o * pr.ldo + c0 + c] = g * s;
        }
    }
    if (pr.bias != nullptr && c0 == 0 && blockIdx.y == 0 && threadIdx.x < 64) {
        const float* bp = a.ws + (long long)a.n_tiles * a.splits * TE + (long long)tile * a.splits * 64 + threadIdx.x;
        float s = 0.f;
        for (int z = 0; z < a.splits; ++z) s += bp[z * 64];
        pr.bias[threadIdx.x] = g * s;
    }
}

int wg_tiles(int n, const elimrec_wgrad_problem_t* p) {
    int t = 0;
    for (int i = 0; i < n; ++i) t += (int)((p[i].K + WG_TN - 1) / WG_TN);
    return t;
}

}  // namespace

ELIMREC_API int64_t elimrec_wgrad_multi_workspace_floats(int n, const elimrec_wgrad_problem_t* problems, int splits) {
    if (n <= 0 || problems == nullptr || splits <= 0) return 0;
    return (int64_t)wg_tiles(n, problems) * splits * (64 * WG_TN + 64);
}

namespace {
int wg_args(const char* who, int n, const elimrec_wgrad_problem_t* problems, int splits, float* workspace, const float* gscale_dev,
            WgArgs* out) {
    WgArgs& a = *out;
    a = WgArgs{};
    a.n_prob = n;
    int tiles = 0;
    for (int i = 0; i < n; ++i) {
        const elimrec_wgrad_problem_t& q = problems[i];
        if (!(q.K > 0 && q.row_begin >= 0 && q.row_end >= q.row_begin)) {
            elimrec_set_error("%s: bad problem shape", who);
            return -1;
        }
        if (!(q.lda % 4 == 0 && (reinterpret_cast<unsigned long long>(q.A) & 15) == 0)) {
            elimrec_set_error("%s: A must be 16-byte aligned rows", who);
            return -1;
        }
        a.p[i] = WgProblem{q.A, q.lda, q.B, q.ldb, (int)q.K, (int)q.row_begin, (int)q.row_end, q.out, q.ldo, q.bias_out,
                           q.scale_by_g, tiles};
        tiles += (int)((q.K + WG_TN - 1) / WG_TN);
    }
    a.n_tiles = tiles;
    a.splits = splits;
    a.ws = workspace;
    a.gscale = gscale_dev;
    return 0;
}
}  // namespace

// the fixed-order reduction of the partial tiles, shared with the tensor-core twin (csrc/linear_tc.cu: elimrec_wgrad_multi_x3)
int er_wgrad_multi_reduce(int n, const elimrec_wgrad_problem_t* problems, int splits, float* workspace, const float* gscale_dev,
                          cudaStream_t st) {
    WgArgs a;
    int rc = wg_args("elimrec_wgrad_multi_x3", n, problems, splits, workspace, gscale_dev, &a);
    if (rc != 0) return rc;
    wgrad_multi_reduce_kernel<<<dim3(a.n_tiles, 32), 256, 0, st>>>(a);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_wgrad_multi(int n, const elimrec_wgrad_problem_t* problems, int splits, float* workspace,
                                    const float* gscale_dev, elimrec_stream_t stream) {
    ER_CHECK_ARG(n >= 0 && n <= ELIMREC_WGRAD_MAX_PROBLEMS && (n == 0 || problems != nullptr), "too many problems");
    ER_CHECK_ARG(splits >= 1 && splits <= 64 && workspace != nullptr, "splits in [1, 64] and a workspace required");
    if (n == 0) return 0;
    WgArgs a;
    int rc = wg_args("elimrec_wgrad_multi", n, problems, splits, workspace, gscale_dev, &a);
    if (rc != 0) return rc;
    cudaStream_t st = er_stream(stream);
    wgrad_multi_kernel<<<dim3(a.n_tiles, splits), 128, 0, st>>>(a);
    ER_LAUNCH_CHECK();
    wgrad_multi_reduce_kernel<<<dim3(a.n_tiles, 32), 256, 0, st>>>(a);
    ER_LAUNCH_CHECK();
    return 0;
}
