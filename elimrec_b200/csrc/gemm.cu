// fp32 (FFMA) strided GEMM + column sums.  Exact-fp32 path for every nn.Linear of the reference
// (models/EliMRec.py:233-236, 146-151, 261-270) and their weight/input gradients; the tensor-core
// (tcgen05 TF32) path for the large, HBM-bound projection layers lives in linear_tc.cu.
//
// C(m,n) = [C +] scale * sum_k A(m,k) B(k,n) [+ bias(n)], element strides for all three operands,
// 128x64x16 tiles, 256 threads, 8x4 register micro-tile, register-staged double buffering.
// split_k > 1: slices write to workspace and a second kernel reduces them in slice order.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, PAD = 4;

struct GemmArgs {
    long long M, N, K;
    const float* A; long long a_sm, a_sk;
    const float* B; long long b_sk, b_sn;
    float* C; long long c_sm, c_sn;
    const float* bias;
    int accumulate;
    int split_k;
    long long k_chunk;
    float* ws;
    const float* scale_dev;
};

__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs g) {
    __shared__ float As[BK][BM + PAD];
    __shared__ float Bs[BK][BN + PAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.y * BM;
    const long long n0 = (long long)blockIdx.x * BN;
    const long long kbeg = (long long)blockIdx.z * g.k_chunk;
    const long long kend = min(g.K, kbeg + g.k_chunk);

    const bool a_kfast = (g.a_sk == 1);
    const bool b_kfast = (g.b_sk == 1 && g.b_sn != 1);
    float ra[8], rb[4];
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    auto load = [&](long long k0) {
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            int mm, kk;
            if (a_kfast) { kk = tid & 15; mm = (tid >> 4) + 16 * p; }
            else { mm = tid & 127; kk = (tid >> 7) + 2 * p; }
            const long long m = m0 + mm, k = k0 + kk;
            ra[p] = (m < g.M && k < kend) ? __ldg(g.A + m * g.a_sm + k * g.a_sk) : 0.f;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            int nn, kk;
            if (b_kfast) { kk = tid & 15; nn = (tid >> 4) + 16 * p; }
            else { nn = tid & 63; kk = (tid >> 6) + 4 * p; }
            const long long n = n0 + nn, k = k0 + kk;
            rb[p] = (n < g.N && k < kend) ? __ldg(g.B + k * g.b_sk + n * g.b_sn) : 0.f;
        }
    };
    auto store = [&]() {
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            int mm, kk;
            if (a_kfast) { kk = tid & 15; mm = (tid >> 4) + 16 * p; }
            else { mm = tid & 127; kk = (tid >> 7) + 2 * p; }
            As[kk][mm] = ra[p];
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            int nn, kk;
            if (b_kfast) { kk = tid & 15; nn = (tid >> 4) + 16 * p; }
            else { nn = tid & 63; kk = (tid >> 6) + 4 * p; }
            Bs[kk][nn] = rb[p];
        }
    };

    if (kbeg < kend) {
        load(kbeg);
        store();
        __syncthreads();
        for (long long k0 = kbeg; k0 < kend; k0 += BK) {
            const bool more = (k0 + BK) < kend;
            if (more) load(k0 + BK);
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
            if (more) {
                store();
                __syncthreads();
            }
        }
    }

    const float scale = (g.scale_dev != nullptr) ? __ldg(g.scale_dev) : 1.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + ty * 8 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            if (g.split_k > 1) {
                g.ws[((long long)blockIdx.z * g.M + m) * g.N + n] = acc[i][j];
            } else {
                float v = acc[i][j] * scale;
                if (g.bias != nullptr) v += __ldg(g.bias + n);
                float* c = g.C + m * g.c_sm + n * g.c_sn;
                if (g.accumulate) v += *c;
                *c = v;
            }
        }
    }
}

__global__ void gemm_splitk_reduce_kernel(GemmArgs g) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.M * g.N) return;
    const long long m = idx / g.N, n = idx % g.N;
    float s = 0.f;
    for (int z = 0; z < g.split_k; ++z) s += g.ws[((long long)z * g.M + m) * g.N + n];
    const float scale = (g.scale_dev != nullptr) ? __ldg(g.scale_dev) : 1.f;
    s *= scale;
    if (g.bias != nullptr) s += __ldg(g.bias + n);
    float* c = g.C + m * g.c_sm + n * g.c_sn;
    if (g.accumulate) s += *c;
    *c = s;
}

constexpr int CS_ROWS = 128;  // rows per colsum chunk

__global__ void colsum_partial_kernel(long long M, long long N, const float* __restrict__ A, long long ld,
                                      float* __restrict__ ws) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long n = (long long)blockIdx.x * 32 + tx;
    const long long r0 = (long long)blockIdx.y * CS_ROWS;
    const long long r1 = min(M, r0 + CS_ROWS);
    float s = 0.f;
    if (n < N) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};   // 4 independent loads in flight per thread
        long long r = r0 + ty;
        for (; r + 24 < r1; r += 32) {
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] += __ldg(A + (r + 8 * u) * ld + n);
        }
        for (; r < r1; r += 8) t[0] += __ldg(A + r * ld + n);
        s = (t[0] + t[1]) + (t[2] + t[3]);
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][tx];
        ws[(long long)blockIdx.y * N + n] = t;
    }
}

// 32 columns x 32 chunk-lanes per CTA: lane (ty) adds chunks ty, ty+32, ... (4 loads in flight), then the 32 partial
// sums of a column are added in a fixed order - deterministic, and ~20 dependent steps instead of n_chunks
__global__ void __launch_bounds__(1024) colsum_final_kernel(long long n_chunks, long long N, const float* __restrict__ ws,
                                                            float* out, int accumulate, const float* scale_dev) {
    __shared__ float red[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long n = (long long)blockIdx.x * 32 + tx;
    float s = 0.f;
    if (n < N) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        long long c = ty;
        for (; c + 96 < n_chunks; c += 128) {
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] += __ldg(ws + (c + 32 * u) * N + n);
        }
        for (; c < n_chunks; c += 32) t[0] += __ldg(ws + c * N + n);
        s = (t[0] + t[1]) + (t[2] + t[3]);
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) t += red[i][tx];
        if (scale_dev != nullptr) t *= __ldg(scale_dev);
        if (accumulate) t += out[n];
        out[n] = t;
    }
}

}  // namespace

ELIMREC_API int64_t elimrec_gemm_workspace_floats(int64_t M, int64_t N, int split_k) {
    return split_k > 1 ? (int64_t)split_k * M * N : 0;
}

ELIMREC_API int elimrec_gemm(int64_t M, int64_t N, int64_t K, const float* A, int64_t a_sm, int64_t a_sk,
                             const float* B, int64_t b_sk, int64_t b_sn, float* C, int64_t c_sm, int64_t c_sn,
                             const float* bias, int accumulate, int split_k, float* workspace, const float* scale_dev,
                             elimrec_stream_t stream) {
    ER_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "negative dimension");
    if (M == 0 || N == 0) return 0;
    if (split_k < 1) split_k = 1;
    ER_CHECK_ARG(split_k == 1 || workspace != nullptr, "split_k > 1 needs a workspace");
    GemmArgs g;
    g.M = M; g.N = N; g.K = K;
    g.A = A; g.a_sm = a_sm; g.a_sk = a_sk;
    g.B = B; g.b_sk = b_sk; g.b_sn = b_sn;
    g.C = C; g.c_sm = c_sm; g.c_sn = c_sn;
    g.bias = bias; g.accumulate = accumulate;
    long long chunk = (K + split_k - 1) / split_k;
    chunk = ((chunk + BK - 1) / BK) * BK;
    if (chunk == 0) chunk = BK;
    split_k = (int)((K + chunk - 1) / chunk);
    if (split_k < 1) split_k = 1;
    g.split_k = split_k; g.k_chunk = chunk; g.ws = workspace; g.scale_dev = scale_dev;
    const long long gy = (M + BM - 1) / BM, gx = (N + BN - 1) / BN;
    ER_CHECK_ARG(gy <= 65535 * 1LL && split_k <= 65535, "grid too large");
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)split_k);
    cudaStream_t st = er_stream(stream);
    gemm_kernel<<<grid, 256, 0, st>>>(g);
    ER_LAUNCH_CHECK();
    if (split_k > 1) {
        const long long tot = M * N;
        gemm_splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g);
        ER_LAUNCH_CHECK();
    }
    return 0;
}

ELIMREC_API int64_t elimrec_colsum_workspace_floats(int64_t M, int64_t N) {
    return ((M + CS_ROWS - 1) / CS_ROWS) * N;
}

ELIMREC_API int elimrec_colsum(int64_t M, int64_t N, const float* A, int64_t ld, float* out, float* workspace,
                               int accumulate, const float* scale_dev, elimrec_stream_t stream) {
    ER_CHECK_ARG(M >= 0 && N > 0, "bad shape");
    ER_CHECK_ARG(workspace != nullptr, "workspace required");
    const long long chunks = (M + CS_ROWS - 1) / CS_ROWS;
    cudaStream_t st = er_stream(stream);
    if (chunks > 0) {
        ER_CHECK_ARG(chunks <= 65535, "too many rows");
        dim3 grid((unsigned)((N + 31) / 32), (unsigned)chunks);
        colsum_partial_kernel<<<grid, 256, 0, st>>>(M, N, A, ld, workspace);
        ER_LAUNCH_CHECK();
    }
    colsum_final_kernel<<<(unsigned)((N + 31) / 32), 1024, 0, st>>>(chunks, N, workspace, out, accumulate, scale_dev);
    ER_LAUNCH_CHECK();
    return 0;
}
