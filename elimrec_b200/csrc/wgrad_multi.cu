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
// of `splits` row ranges (register-staged double buffering, 4 x 4 micro-tile), partial tiles go to scratch and are summed in
// split order.
#include "common.cuh"

namespace {

constexpr int WG_BR = 16;          // rows per staged chunk

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
    float* ws;          // [n_tiles][splits][64 x 64] partial tiles, then [n_tiles][splits][64] partial column sums
    const float* gscale;
};

__device__ __forceinline__ int wg_find(const WgArgs& a, int tile) {
    int q = 0;
    while (q + 1 < a.n_prob && tile >= a.p[q + 1].tile0) ++q;
    return q;
}

__global__ void __launch_bounds__(256) wgrad_multi_kernel(const __grid_constant__ WgArgs a) {
    __shared__ float As[2][WG_BR][64 + 4];
    __shared__ float Bs[2][WG_BR][64 + 4];
    const int tile = blockIdx.x, split = blockIdx.y;
    const WgProblem& pr = a.p[wg_find(a, tile)];
    const int c0 = (tile - pr.tile0) * 64;                 // first column of B / out handled here
    const int rows = pr.r1 - pr.r0;
    int chunk = (rows + a.splits - 1) / a.splits;
    chunk = (chunk + WG_BR - 1) / WG_BR * WG_BR;
    const int rb = pr.r0 + split * chunk, re = min(pr.r1, rb + chunk);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lr = tid >> 4, lc = (tid & 15) * 4;          // this thread's float4 of a staged chunk: row lr, columns lc..lc+3
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;
    const bool want_bias = pr.bias != nullptr && c0 == 0;
    const bool b_vec = (pr.ldb % 4 == 0) && ((reinterpret_cast<unsigned long long>(pr.B) & 15) == 0) && (c0 + 64 <= pr.K);

    float4 ra, rbv;
    auto load = [&](int r) {
        const int row = r + lr;
        ra = make_float4(0.f, 0.f, 0.f, 0.f);
        rbv = ra;
        if (row < re) {
            ra = __ldg(reinterpret_cast<const float4*>(pr.A + (long long)row * pr.lda + lc));
            const float* bp = pr.B + (long long)row * pr.ldb + c0 + lc;
            if (b_vec) {
                rbv = __ldg(reinterpret_cast<const float4*>(bp));
            } else {
                const int left = pr.K - (c0 + lc);
                if (left > 0) rbv.x = __ldg(bp);
                if (left > 1) rbv.y = __ldg(bp + 1);
                if (left > 2) rbv.z = __ldg(bp + 2);
                if (left > 3) rbv.w = __ldg(bp + 3);
            }
        }
    };
    auto store = [&](int buf) {
        *reinterpret_cast<float4*>(&As[buf][lr][lc]) = ra;
        *reinterpret_cast<float4*>(&Bs[buf][lr][lc]) = rbv;
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
                const float4 av = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
                const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                const float a4[4] = {av.x, av.y, av.z, av.w};
                const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
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
    float* part = a.ws + ((long long)tile * a.splits + split) * 4096;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(part + (ty * 4 + i) * 64 + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (want_bias && tid < 64) a.ws[(long long)a.n_tiles * a.splits * 4096 + ((long long)tile * a.splits + split) * 64 + tid] = bsum;
}

// grid (tiles, 16): CTA y sums 256 elements of the tile over the splits, in split order
__global__ void __launch_bounds__(256) wgrad_multi_reduce_kernel(const __grid_constant__ WgArgs a) {
    const int tile = blockIdx.x;
    const WgProblem& pr = a.p[wg_find(a, tile)];
    const int c0 = (tile - pr.tile0) * 64;
    const float g = (pr.scale_by_g && a.gscale != nullptr) ? __ldg(a.gscale) : 1.f;
    const float* part = a.ws + (long long)tile * a.splits * 4096;
    {
        const int e = blockIdx.y * 256 + threadIdx.x;
        const int o = e >> 6, c = e & 63;
        if (c0 + c < pr.K) {
            float s = 0.f;
            int z = 0;
            for (; z + 4 <= a.splits; z += 4) {
                const float t0 = part[(long long)z * 4096 + e], t1 = part[(long long)(z + 1) * 4096 + e];
                const float t2 = part[(long long)(z + 2) * 4096 + e], t3 = part[(long long)(z + 3) * 4096 + e];
                s = (((s + t0) + t1) + t2) + t3;
            }
            for (; z < a.splits; ++z) s += part[(long long)z * 4096 + e];
            pr.out[(long long)o * pr.ldo + c0 + c] = g * s;
        }
    }
    if (pr.bias != nullptr && c0 == 0 && blockIdx.y == 0 && threadIdx.x < 64) {
        const float* bp = a.ws + (long long)a.n_tiles * a.splits * 4096 + (long long)tile * a.splits * 64 + threadIdx.x;
        float s = 0.f;
        for (int z = 0; z < a.splits; ++z) s += bp[z * 64];
        pr.bias[threadIdx.x] = g * s;
    }
}

int wg_tiles(int n, const elimrec_wgrad_problem_t* p) {
    int t = 0;
    for (int i = 0; i < n; ++i) t += (int)((p[i].K + 63) / 64);
    return t;
}

}  // namespace

ELIMREC_API int64_t elimrec_wgrad_multi_workspace_floats(int n, const elimrec_wgrad_problem_t* problems, int splits) {
    if (n <= 0 || problems == nullptr || splits <= 0) return 0;
    return (int64_t)wg_tiles(n, problems) * splits * (4096 + 64);
}

ELIMREC_API int elimrec_wgrad_multi(int n, const elimrec_wgrad_problem_t* problems, int splits, float* workspace,
                                    const float* gscale_dev, elimrec_stream_t stream) {
    ER_CHECK_ARG(n >= 0 && n <= ELIMREC_WGRAD_MAX_PROBLEMS && (n == 0 || problems != nullptr), "too many problems");
    ER_CHECK_ARG(splits >= 1 && splits <= 64 && workspace != nullptr, "splits in [1, 64] and a workspace required");
    if (n == 0) return 0;
    WgArgs a{};
    a.n_prob = n;
    int tiles = 0;
    for (int i = 0; i < n; ++i) {
        const elimrec_wgrad_problem_t& q = problems[i];
        ER_CHECK_ARG(q.K > 0 && q.row_begin >= 0 && q.row_end >= q.row_begin, "bad problem shape");
        ER_CHECK_ARG(q.lda % 4 == 0 && (reinterpret_cast<unsigned long long>(q.A) & 15) == 0, "A must be 16-byte aligned rows");
        a.p[i] = WgProblem{q.A, q.lda, q.B, q.ldb, (int)q.K, (int)q.row_begin, (int)q.row_end, q.out, q.ldo, q.bias_out,
                           q.scale_by_g, tiles};
        tiles += (int)((q.K + 63) / 64);
    }
    a.n_tiles = tiles;
    a.splits = splits;
    a.ws = workspace;
    a.gscale = gscale_dev;
    cudaStream_t st = er_stream(stream);
    wgrad_multi_kernel<<<dim3(tiles, splits), 256, 0, st>>>(a);
    ER_LAUNCH_CHECK();
    wgrad_multi_reduce_kernel<<<dim3(tiles, 16), 256, 0, st>>>(a);
    ER_LAUNCH_CHECK();
    return 0;
}
