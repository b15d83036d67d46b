// Fused BPR step: row gathers + F.normalize + cosine scores + softplus + mean for the fused table and
// every single-modal head, forward AND backward in one pass (reference models/EliMRec.py:115-142,
// 277-297: getEmbedding gathers, original_bpr_loss x (1+M), and their autograd graph of ~40 kernels).
//
// One warp per (triple k, table t); a lane owns 2 of the 64 dimensions.  Five warp reductions give
// |u|^2 |p|^2 |n|^2 and the two cosines; the gradient of the normalised dot products is closed-form
// (no extra reductions).  The gradient is emitted row-sparse: 3B "instance" rows (u | U+pos | U+neg)
// x (64 columns per table), already weighted by weight[t]/B; the dense [N x 64] table gradients of
// the reference are never materialised.
#include "common.cuh"

namespace {

struct BprTables {
    const float* t[1 + ELIMREC_MAX_MODS];
    float w[1 + ELIMREC_MAX_MODS];
};

__device__ __forceinline__ float softplus_ref(float x) {  // F.softplus(beta=1, threshold=20)
    return x > 20.f ? x : log1pf(expf(x));
}

__global__ void __launch_bounds__(256)
bpr_kernel(int B, int n_tables, BprTables tb, const long long* __restrict__ users, const long long* __restrict__ pos,
           const long long* __restrict__ neg, int num_users, int* __restrict__ inst_rows, float* __restrict__ inst_grad,
           float* __restrict__ terms) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= B * n_tables) return;
    const int t = w / B, k = w % B;
    const int u = (int)users[k], p = (int)pos[k] + num_users, n = (int)neg[k] + num_users;
    const float* T = tb.t[t];
    const float2 uv = __ldg(reinterpret_cast<const float2*>(T + (long long)u * 64) + lane);
    const float2 pv = __ldg(reinterpret_cast<const float2*>(T + (long long)p * 64) + lane);
    const float2 nv = __ldg(reinterpret_cast<const float2*>(T + (long long)n * 64) + lane);
    const float eps = 1e-12f;
    const float cu = fmaxf(sqrtf(warp_sum(uv.x * uv.x + uv.y * uv.y)), eps);
    const float cp = fmaxf(sqrtf(warp_sum(pv.x * pv.x + pv.y * pv.y)), eps);
    const float cn = fmaxf(sqrtf(warp_sum(nv.x * nv.x + nv.y * nv.y)), eps);
    const float2 uh = make_float2(uv.x / cu, uv.y / cu);
    const float2 ph = make_float2(pv.x / cp, pv.y / cp);
    const float2 nh = make_float2(nv.x / cn, nv.y / cn);
    const float ps = warp_sum(uh.x * ph.x + uh.y * ph.y);
    const float ns = warp_sum(uh.x * nh.x + uh.y * nh.y);
    const float x = ns - ps;
    if (lane == 0) terms[t * B + k] = softplus_ref(x);
    // d softplus = sigmoid(x) (x > threshold: 1)
    const float sg = x > 20.f ? 1.f : 1.f / (1.f + expf(-x));
    const float gs = tb.w[t] * sg / (float)B;
    // d/du = (gs (n^ - p^) - u^ gs (ns - ps)) / |u| ; d/dn = gs (u^ - n^ ns) / |n| ; d/dp = -gs (u^ - p^ ps) / |p|
    float2 du, dp, dn;
    du.x = (gs * (nh.x - ph.x) - uh.x * gs * x) / cu;
    du.y = (gs * (nh.y - ph.y) - uh.y * gs * x) / cu;
    dn.x = gs * (uh.x - nh.x * ns) / cn;
    dn.y = gs * (uh.y - nh.y * ns) / cn;
    dp.x = -gs * (uh.x - ph.x * ps) / cp;
    dp.y = -gs * (uh.y - ph.y * ps) / cp;
    const long long ld = 64LL * n_tables;
    reinterpret_cast<float2*>(inst_grad + (long long)k * ld + t * 64)[lane] = du;
    reinterpret_cast<float2*>(inst_grad + (long long)(B + k) * ld + t * 64)[lane] = dp;
    reinterpret_cast<float2*>(inst_grad + (long long)(2 * B + k) * ld + t * 64)[lane] = dn;
    if (t == 0 && lane == 0) {
        inst_rows[k] = u;
        inst_rows[B + k] = p;
        inst_rows[2 * B + k] = n;
    }
}

// loss = sum_t w[t] * mean_k terms[t][k]; single block, fixed order => deterministic.
__global__ void bpr_loss_reduce_kernel(int B, int n_tables, BprTables tb, const float* __restrict__ terms,
                                       float* __restrict__ loss) {
    __shared__ float red[256];
    float total = 0.f;
    for (int t = 0; t < n_tables; ++t) {
        float s = 0.f;
        for (int k = threadIdx.x; k < B; k += 256) s += terms[t * B + k];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) total += tb.w[t] * (red[0] / (float)B);
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = total;
}

__global__ void adam_tick_kernel(long long* step, double* consts, double lr, double b1, double b2) {
    const long long t = *step + 1;
    *step = t;
    const double bc1 = 1.0 - pow(b1, (double)t);
    const double bc2 = 1.0 - pow(b2, (double)t);
    consts[0] = lr / bc1;      // step_size
    consts[1] = sqrt(bc2);     // bias_correction2_sqrt
}

// torch.optim.Adam single-tensor math (torch/optim/adam.py, non-amsgrad, maximize=False):
//   g += wd * p;  m.lerp_(g, 1-b1);  v = v*b2 + (1-b2) g*g;  p -= step_size * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void adam_apply_kernel(long long n, float* __restrict__ p, const float* __restrict__ g, long long row_len,
                                  long long g_ld, float* __restrict__ m, float* __restrict__ v,
                                  const double* __restrict__ consts, double b1d, double b2d, float eps, float wd) {
    const float step_size = (float)consts[0];
    const float bc2s = (float)consts[1];
    const float b2 = (float)b2d;
    const float omb1 = (float)(1.0 - b1d), omb2 = (float)(1.0 - b2d);  // python doubles rounded once, as torch does
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long gi = (g_ld == row_len) ? i : (i / row_len) * g_ld + (i % row_len);
        const float pi = p[i];
        float gr = g[gi];
        gr = fmaf(wd, pi, gr);
        float mi = m[i];
        mi = mi + (gr - mi) * omb1;
        float vi = v[i] * b2;
        vi = fmaf(omb2 * gr, gr, vi);
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2s + eps;
        p[i] = pi - step_size * (mi / denom);
    }
}

}  // namespace

ELIMREC_API int elimrec_bpr_forward_backward(int B, int n_tables, const float* const* tables_host,
                                             const float* weight_host, const int64_t* users, const int64_t* pos,
                                             const int64_t* neg, int32_t num_users, float* loss_out,
                                             int32_t* inst_rows, float* inst_grad, float* workspace,
                                             elimrec_stream_t stream) {
    ER_CHECK_ARG(B > 0, "empty batch");
    ER_CHECK_ARG(n_tables >= 1 && n_tables <= 1 + ELIMREC_MAX_MODS, "n_tables out of range");
    BprTables tb{};
    for (int t = 0; t < n_tables; ++t) {
        tb.t[t] = tables_host[t];
        tb.w[t] = weight_host[t];
    }
    cudaStream_t st = er_stream(stream);
    const long long warps = (long long)B * n_tables;
    bpr_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(B, n_tables, tb, (const long long*)users, (const long long*)pos,
                                                             (const long long*)neg, num_users, inst_rows, inst_grad,
                                                             workspace);
    ER_LAUNCH_CHECK();
    bpr_loss_reduce_kernel<<<1, 256, 0, st>>>(B, n_tables, tb, workspace, loss_out);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_adam_tick(int64_t* step_dev, double* consts_dev, double lr, double beta1, double beta2,
                                  elimrec_stream_t stream) {
    adam_tick_kernel<<<1, 1, 0, er_stream(stream)>>>((long long*)step_dev, consts_dev, lr, beta1, beta2);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_adam_apply(int64_t n, float* param, const float* grad, int64_t row_len, int64_t grad_ld,
                                   float* exp_avg, float* exp_avg_sq, const double* consts_dev, double beta1, double beta2,
                                   float eps, float weight_decay, elimrec_stream_t stream) {
    ER_CHECK_ARG(row_len > 0 && grad_ld >= row_len, "bad gradient view");
    if (n <= 0) return 0;
    long long blocks = (n + 1023) / 1024;
    if (blocks > 148 * 16) blocks = 148 * 16;
    adam_apply_kernel<<<(unsigned)blocks, 256, 0, er_stream(stream)>>>(n, param, grad, row_len, grad_ld, exp_avg,
                                                                       exp_avg_sq, consts_dev, beta1, beta2, eps,
                                                                       weight_decay);
    ER_LAUNCH_CHECK();
    return 0;
}
