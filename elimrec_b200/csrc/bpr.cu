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


// ---------------------------------------------------------------------------------------------
// Backward of the fusion Linear (embedding_{user,item}_after_GCN) and the single-modal heads (s_dense_*) on the
// 3B instance rows only (their gradient is zero on every other row).  Three launches replace the ~18 tiny
// GEMM / column-sum launches a generic implementation needs.
//   inst_dO   : dO[r, :] = g * ( dF[r] @ W_{u|i}  +  [0 | dS_v[r] @ Ws_v | dS_a[r] @ Ws_a | dS_t[r] @ Ws_t] )
//   inst_dW   : per 64-row chunk, partial outer products  dY^T O  for every weight (and column sums for the biases)
//   inst_dWred: fixed-order sum of the chunk partials (deterministic), scaled by the upstream gradient g
// ---------------------------------------------------------------------------------------------
struct InstW {
    const float* Wu;   // [64 x F]
    const float* Wi;   // [64 x F]
    const float* Ws[ELIMREC_MAX_MODS];  // [64 x 64]
};

constexpr int IRB = 16;  // rows per CTA in inst_dO

// backward seeds of the linear schedule fused into inst_dO (what elimrec_lin_seed2 does in a launch of its own):
//   GA[rows[j]] += scale * (sum of all 64-column blocks of dO[j]),  GB[rows[j]] += scale * dO[j, 0:64]
struct InstSeed {
    const int* rows;    // NULL: no seeds
    float* GA;
    float* GB;
    long long ldg;
    float scale;
    int n_mod;
};

__global__ void __launch_bounds__(256)
inst_dO_kernel(int B, int nt, int F, InstW w, const float* __restrict__ ig, const float* __restrict__ gscale,
               float* __restrict__ dO, InstSeed sd) {
    __shared__ __align__(16) float gs[IRB][64 * (1 + ELIMREC_MAX_MODS)];
    const int nbu = (B + IRB - 1) / IRB;
    const bool user = (int)blockIdx.x < nbu;
    const int r0 = user ? blockIdx.x * IRB : B + (blockIdx.x - nbu) * IRB;
    const int r1 = min(user ? B : 3 * B, r0 + IRB);
    const int ld = 64 * nt;
    for (int i = threadIdx.x; i < IRB * ld; i += blockDim.x) {
        const int r = i / ld, c = i % ld;
        gs[r][c] = (r0 + r < r1) ? ig[(long long)(r0 + r) * ld + c] : 0.f;
    }
    __syncthreads();
    const float g = (gscale != nullptr) ? __ldg(gscale) : 1.f;
    const float* W = user ? w.Wu : w.Wi;
    // the instance gradients come out of shared memory four at a time (one broadcast LDS.128 per 4 FMAs)
    for (int c = threadIdx.x; c < F; c += blockDim.x) {
        float acc[IRB];
#pragma unroll
        for (int r = 0; r < IRB; ++r) acc[r] = 0.f;
        for (int n = 0; n < 64; n += 4) {
            const float w0 = __ldg(W + n * F + c), w1 = __ldg(W + (n + 1) * F + c);
            const float w2 = __ldg(W + (n + 2) * F + c), w3 = __ldg(W + (n + 3) * F + c);
#pragma unroll
            for (int r = 0; r < IRB; ++r) {
                const float4 gv = *reinterpret_cast<const float4*>(&gs[r][n]);
                acc[r] = fmaf(gv.w, w3, fmaf(gv.z, w2, fmaf(gv.y, w1, fmaf(gv.x, w0, acc[r]))));
            }
        }
        const int blk = c >> 6;
        if (blk >= 1) {
            const float* Ws = w.Ws[blk - 1];
            const int cc = c & 63;
            for (int n = 0; n < 64; n += 4) {
                const float w0 = __ldg(Ws + n * 64 + cc), w1 = __ldg(Ws + (n + 1) * 64 + cc);
                const float w2 = __ldg(Ws + (n + 2) * 64 + cc), w3 = __ldg(Ws + (n + 3) * 64 + cc);
#pragma unroll
                for (int r = 0; r < IRB; ++r) {
                    const float4 gv = *reinterpret_cast<const float4*>(&gs[r][blk * 64 + n]);
                    acc[r] = fmaf(gv.w, w3, fmaf(gv.z, w2, fmaf(gv.y, w1, fmaf(gv.x, w0, acc[r]))));
                }
            }
        }
#pragma unroll
        for (int r = 0; r < IRB; ++r)
            if (r0 + r < r1) dO[(long long)(r0 + r) * F + c] = g * acc[r];
        if (sd.rows != nullptr) {      // the tile goes back through shared memory for the cross-block sums of the seeds
            __syncthreads();           // (F = blockDim.x columns: every thread is here exactly once; gs is no longer read)
#pragma unroll
            for (int r = 0; r < IRB; ++r) gs[r][c] = g * acc[r];
        }
    }
    if (sd.rows != nullptr) {
        __syncthreads();
        for (int i = threadIdx.x; i < IRB * 64; i += blockDim.x) {
            const int r = i >> 6, col = i & 63;
            if (r0 + r >= r1) continue;
            const float b = gs[r][col];
            float a = b;
            for (int m = 0; m < sd.n_mod; ++m) a += gs[r][64 * (m + 1) + col];
            const long long node = __ldg(sd.rows + r0 + r);
            atomicAdd(sd.GA + node * sd.ldg + col, sd.scale * a);
            atomicAdd(sd.GB + node * sd.ldg + col, sd.scale * b);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Forward twin of inst_dO on the 3B instance rows (row-sparse step): fusion Linear + single-modal heads of
//   F[r]   = O[r, 0:F] @ W_{u|i}^T + b_{u|i}          (models/EliMRec.py:261-270, rows 0..B-1 are users)
//   S_m[r] = O[r, block m] @ Ws_m^T + bs_m            (models/EliMRec.py:146-151)
// in exact fp32 FFMA.  A CTA takes 16 rows (staged in shared memory, read back as broadcast LDS.128); thread (q, col) owns
// output column col and the 64-wide K block q: the K-block-q partial of the fusion Linear and, for q >= 1, head q-1.  The
// nt fusion partials of an output are added in block order through shared memory (deterministic).  Measured at Tiktok
// shape: 29 us (instruction-issue bound: ~3000 instructions per warp on 2-3 CTAs per SM) against 21 us for the two
// tensor-core launches (fuse_heads_x3, users beside items) - the step therefore keeps the tensor-core form and this kernel
// is the exact-fp32 option (model config inst_fuse="ffma").
// ---------------------------------------------------------------------------------------------
struct InstFwd {
    const float* Wu;   // [64 x F]
    const float* Wi;
    const float* Ws[ELIMREC_MAX_MODS];  // [64 x 64]
    const float* bu;
    const float* bi;
    const float* bs[ELIMREC_MAX_MODS];
    float* Fout;                        // [3B x 64]
    float* Sout[ELIMREC_MAX_MODS];      // [3B x 64]
};

__global__ void __launch_bounds__(64 * (1 + ELIMREC_MAX_MODS))
inst_fwd_kernel(int B, int nt, int F, InstFwd w, const float* __restrict__ Oin) {
    __shared__ __align__(16) float os[IRB][64 * (1 + ELIMREC_MAX_MODS)];
    __shared__ float red[1 + ELIMREC_MAX_MODS][IRB][64];
    const int nbu = (B + IRB - 1) / IRB;
    const bool user = (int)blockIdx.x < nbu;
    const int r0 = user ? blockIdx.x * IRB : B + (blockIdx.x - nbu) * IRB;
    const int r1 = min(user ? B : 3 * B, r0 + IRB);
    for (int i = threadIdx.x; i < IRB * (F / 4); i += blockDim.x) {
        const int r = i / (F / 4), c4 = i % (F / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < r1) v = __ldg(reinterpret_cast<const float4*>(Oin + (long long)(r0 + r) * F) + c4);
        *reinterpret_cast<float4*>(&os[r][c4 * 4]) = v;
    }
    __syncthreads();
    const int q = threadIdx.x >> 6, col = threadIdx.x & 63;
    const float4* Wf = reinterpret_cast<const float4*>((user ? w.Wu : w.Wi) + (long long)col * F + q * 64);
    float accf[IRB], acch[IRB];
#pragma unroll
    for (int r = 0; r < IRB; ++r) { accf[r] = 0.f; acch[r] = 0.f; }
    for (int k4 = 0; k4 < 16; ++k4) {
        const float4 wv = __ldg(Wf + k4);
#pragma unroll
        for (int r = 0; r < IRB; ++r) {
            const float4 o = *reinterpret_cast<const float4*>(&os[r][q * 64 + k4 * 4]);
            accf[r] = fmaf(o.w, wv.w, fmaf(o.z, wv.z, fmaf(o.y, wv.y, fmaf(o.x, wv.x, accf[r]))));
        }
    }
    if (q >= 1) {
        const float4* Wh = reinterpret_cast<const float4*>(w.Ws[q - 1] + col * 64);
        for (int k4 = 0; k4 < 16; ++k4) {
            const float4 wv = __ldg(Wh + k4);
#pragma unroll
            for (int r = 0; r < IRB; ++r) {
                const float4 o = *reinterpret_cast<const float4*>(&os[r][q * 64 + k4 * 4]);
                acch[r] = fmaf(o.w, wv.w, fmaf(o.z, wv.z, fmaf(o.y, wv.y, fmaf(o.x, wv.x, acch[r]))));
            }
        }
        const float b = __ldg(w.bs[q - 1] + col);
        float* So = w.Sout[q - 1];
#pragma unroll
        for (int r = 0; r < IRB; ++r)
            if (r0 + r < r1) So[(long long)(r0 + r) * 64 + col] = acch[r] + b;
    }
#pragma unroll
    for (int r = 0; r < IRB; ++r) red[q][r][col] = accf[r];
    __syncthreads();
    const float* bias = user ? w.bu : w.bi;
    for (int i = threadIdx.x; i < IRB * 64; i += blockDim.x) {
        const int r = i >> 6, c = i & 63;
        if (r0 + r < r1) {
            float sum = red[0][r][c];
            for (int t = 1; t < nt; ++t) sum += red[t][r][c];
            w.Fout[(long long)(r0 + r) * 64 + c] = sum + __ldg(bias + c);
        }
    }
}

constexpr int IRC = 32;  // rows per chunk in inst_dW (32 KB of instance gradients in shared memory)

// partial layout per chunk: [64 x F] fusion weight | [M][64 x 64] heads | [64*nt] bias column sums
__global__ void __launch_bounds__(256)
inst_dW_kernel(int B, int nt, int F, const float* __restrict__ ig, const float* __restrict__ Oin, float* __restrict__ part) {
    __shared__ float gs[IRC][64 * (1 + ELIMREC_MAX_MODS)];
    const int ncu = (B + IRC - 1) / IRC;
    const bool user = (int)blockIdx.x < ncu;
    const int r0 = user ? blockIdx.x * IRC : B + (blockIdx.x - ncu) * IRC;
    const int r1 = min(user ? B : 3 * B, r0 + IRC);
    const int nr = r1 - r0;
    const int ld = 64 * nt;
    for (int i = threadIdx.x; i < IRC * ld; i += blockDim.x) {
        const int r = i / ld, c = i % ld;
        gs[r][c] = (r < nr) ? ig[(long long)(r0 + r) * ld + c] : 0.f;
    }
    __syncthreads();
    const long long psz = 64LL * F + (long long)(nt - 1) * 4096 + ld;
    float* out = part + (long long)blockIdx.x * psz;
    for (int c = threadIdx.x; c < F; c += blockDim.x) {
        float acc[64];
#pragma unroll
        for (int n = 0; n < 64; ++n) acc[n] = 0.f;
        for (int r = 0; r < nr; ++r) {
            const float o = __ldg(Oin + (long long)(r0 + r) * F + c);
#pragma unroll
            for (int n = 0; n < 64; ++n) acc[n] = fmaf(gs[r][n], o, acc[n]);
        }
#pragma unroll
        for (int n = 0; n < 64; ++n) out[(long long)n * F + c] = acc[n];
        const int blk = c >> 6;
        if (blk >= 1) {
#pragma unroll
            for (int n = 0; n < 64; ++n) acc[n] = 0.f;
            for (int r = 0; r < nr; ++r) {
                const float o = __ldg(Oin + (long long)(r0 + r) * F + c);
#pragma unroll
                for (int n = 0; n < 64; ++n) acc[n] = fmaf(gs[r][blk * 64 + n], o, acc[n]);
            }
            float* hw = out + 64LL * F + (long long)(blk - 1) * 4096;
#pragma unroll
            for (int n = 0; n < 64; ++n) hw[n * 64 + (c & 63)] = acc[n];
        }
    }
    for (int c = threadIdx.x; c < ld; c += blockDim.x) {
        float sacc = 0.f;
        for (int r = 0; r < nr; ++r) sacc += gs[r][c];
        out[64LL * F + (long long)(nt - 1) * 4096 + c] = sacc;
    }
}

struct InstOut {
    float* dWu; float* dWi; float* dbu; float* dbi;
    float* dWs[ELIMREC_MAX_MODS]; float* dbs[ELIMREC_MAX_MODS];
};

__global__ void inst_dWred_kernel(int B, int nt, int F, const float* __restrict__ part, const float* __restrict__ gscale,
                                  InstOut o) {
    const int ncu = (B + IRC - 1) / IRC, nci = (2 * B + IRC - 1) / IRC;
    const int ld = 64 * nt;
    const long long psz = 64LL * F + (long long)(nt - 1) * 4096 + ld;
    const long long nWf = 64LL * F, nWs = (long long)(nt - 1) * 4096;
    const long long total = 2 * nWf + nWs + 64 * 2 + 64 * (nt - 1);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float g = (gscale != nullptr) ? __ldg(gscale) : 1.f;
    int c0, c1;          // chunk range
    long long off;       // offset inside a chunk partial
    float* dst;
    if (i < nWf) { c0 = 0; c1 = ncu; off = i; dst = o.dWu + i; }
    else if (i < 2 * nWf) { c0 = ncu; c1 = ncu + nci; off = i - nWf; dst = o.dWi + (i - nWf); }
    else if (i < 2 * nWf + nWs) {
        const long long j = i - 2 * nWf;
        c0 = 0; c1 = ncu + nci; off = nWf + j; dst = o.dWs[j / 4096] + (j % 4096);
    } else {
        const long long j = i - 2 * nWf - nWs;   // biases: [bu 64 | bi 64 | bs_m 64 each]
        if (j < 64) { c0 = 0; c1 = ncu; off = nWf + nWs + j; dst = o.dbu + j; }
        else if (j < 128) { c0 = ncu; c1 = ncu + nci; off = nWf + nWs + (j - 64); dst = o.dbi + (j - 64); }
        else { const long long q = j - 128; c0 = 0; c1 = ncu + nci; off = nWf + nWs + 64 + q; dst = o.dbs[q / 64] + (q % 64); }
    }
    float sacc = 0.f;
    for (int c = c0; c < c1; ++c) sacc += part[(long long)c * psz + off];
    *dst = g * sacc;
}

struct AdamMulti {
    int n;
    float* p[ELIMREC_ADAM_MAX_TENSORS];
    const float* g[ELIMREC_ADAM_MAX_TENSORS];
    float* m[ELIMREC_ADAM_MAX_TENSORS];
    float* v[ELIMREC_ADAM_MAX_TENSORS];
    long long numel[ELIMREC_ADAM_MAX_TENSORS];
    long long row_len[ELIMREC_ADAM_MAX_TENSORS];
    long long g_ld[ELIMREC_ADAM_MAX_TENSORS];
    int block_start[ELIMREC_ADAM_MAX_TENSORS + 1];
};
constexpr int ADAM_CHUNK = 2048;  // elements per CTA

__global__ void __launch_bounds__(256)
adam_multi_kernel(const __grid_constant__ AdamMulti a, const double* __restrict__ consts, double b1d, double b2d, float eps,
                  float wd) {
    int t = 0;
    while (t + 1 < a.n && (int)blockIdx.x >= a.block_start[t + 1]) ++t;
    const long long base = (long long)(blockIdx.x - a.block_start[t]) * ADAM_CHUNK;
    const long long end = min(a.numel[t], base + ADAM_CHUNK);
    const float step_size = (float)consts[0];
    const float bc2s = (float)consts[1];
    const float b2 = (float)b2d;
    const float omb1 = (float)(1.0 - b1d), omb2 = (float)(1.0 - b2d);
    float* __restrict__ p = a.p[t];
    const float* __restrict__ g = a.g[t];
    float* __restrict__ m = a.m[t];
    float* __restrict__ v = a.v[t];
    const long long rl = a.row_len[t], gl = a.g_ld[t];
    for (long long i = base + threadIdx.x; i < end; i += 256) {
        const long long gi = (gl == rl) ? i : (i / rl) * gl + (i % rl);
        const float pi = p[i];
        float gr = fmaf(wd, pi, g[gi]);
        float mi = m[i];
        mi = mi + (gr - mi) * omb1;
        float vi = fmaf(omb2 * gr, gr, v[i] * b2);
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - step_size * (mi / (sqrtf(vi) / bc2s + eps));
    }
}

}  // namespace

ELIMREC_API int elimrec_bpr_forward_backward(int B, int n_tables, const float* const* tables_host,
                                             const float* weight_host, const int64_t* users, const int64_t* pos,
                                             const int64_t* neg, int32_t num_users, float* loss_out,
                                             int32_t* inst_rows, float* inst_grad, float* workspace,
                                             elimrec_stream_t stream) {
    return elimrec_bpr_forward_backward_part(3, B, n_tables, tables_host, weight_host, users, pos, neg, num_users, loss_out,
                                             inst_rows, inst_grad, workspace, stream);
}

ELIMREC_API int elimrec_bpr_forward_backward_part(int part, int B, int n_tables, const float* const* tables_host,
                                                  const float* weight_host, const int64_t* users, const int64_t* pos,
                                                  const int64_t* neg, int32_t num_users, float* loss_out,
                                                  int32_t* inst_rows, float* inst_grad, float* workspace,
                                                  elimrec_stream_t stream) {
    ER_CHECK_ARG(part >= 1 && part <= 3, "part is 1 (terms + instance gradients), 2 (loss reduction) or 3 (both)");
    ER_CHECK_ARG(B > 0, "empty batch");
    ER_CHECK_ARG(n_tables >= 1 && n_tables <= 1 + ELIMREC_MAX_MODS, "n_tables out of range");
    BprTables tb{};
    for (int t = 0; t < n_tables; ++t) {
        tb.t[t] = tables_host[t];
        tb.w[t] = weight_host[t];
    }
    cudaStream_t st = er_stream(stream);
    const long long warps = (long long)B * n_tables;
    if (part & 1) {
        bpr_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(B, n_tables, tb, (const long long*)users, (const long long*)pos,
                                                                 (const long long*)neg, num_users, inst_rows, inst_grad,
                                                                 workspace);
        ER_LAUNCH_CHECK();
    }
    if (part & 2) {
        bpr_loss_reduce_kernel<<<1, 256, 0, st>>>(B, n_tables, tb, workspace, loss_out);
        ER_LAUNCH_CHECK();
    }
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


ELIMREC_API int64_t elimrec_inst_backward_workspace_floats(int B, int n_tables, int F) {
    const int64_t chunks = (B + IRC - 1) / IRC + (2 * B + IRC - 1) / IRC;
    return chunks * (64LL * F + (int64_t)(n_tables - 1) * 4096 + 64 * n_tables);
}

ELIMREC_API int elimrec_inst_backward(int B, int n_tables, int F, const float* inst_grad, const float* O_inst,
                                      const float* gscale_dev, const float* Wu, const float* Wi,
                                      const float* const* Ws_host, float* dO_inst, float* dWu, float* dWi, float* dbu,
                                      float* dbi, float* const* dWs_host, float* const* dbs_host, float* workspace,
                                      elimrec_stream_t stream) {
    return elimrec_inst_backward_part(3, B, n_tables, F, inst_grad, O_inst, gscale_dev, Wu, Wi, Ws_host, dO_inst, dWu, dWi, dbu,
                                      dbi, dWs_host, dbs_host, workspace, stream);
}

ELIMREC_API int elimrec_inst_backward_part(int part, int B, int n_tables, int F, const float* inst_grad, const float* O_inst,
                                           const float* gscale_dev, const float* Wu, const float* Wi,
                                           const float* const* Ws_host, float* dO_inst, float* dWu, float* dWi, float* dbu,
                                           float* dbi, float* const* dWs_host, float* const* dbs_host, float* workspace,
                                           elimrec_stream_t stream) {
    ER_CHECK_ARG(part >= 1 && part <= 3, "part must be 1 (dO), 2 (weight gradients) or 3 (both)");
    ER_CHECK_ARG(B > 0 && n_tables >= 1 && n_tables <= 1 + ELIMREC_MAX_MODS, "bad batch / table count");
    ER_CHECK_ARG(F == 64 * n_tables, "F must be 64 * n_tables (concat fusion)");
    InstW w{};
    w.Wu = Wu; w.Wi = Wi;
    InstOut o{};
    o.dWu = dWu; o.dWi = dWi; o.dbu = dbu; o.dbi = dbi;
    for (int m = 0; m < n_tables - 1; ++m) {
        w.Ws[m] = Ws_host[m];
        o.dWs[m] = dWs_host[m];
        o.dbs[m] = dbs_host[m];
    }
    cudaStream_t st = er_stream(stream);
    if (part & 1) {
        const int nb = (B + IRB - 1) / IRB + (2 * B + IRB - 1) / IRB;
        inst_dO_kernel<<<nb, 256, 0, st>>>(B, n_tables, F, w, inst_grad, gscale_dev, dO_inst, InstSeed{});
        ER_LAUNCH_CHECK();
    }
    if (!(part & 2)) return 0;
    const int nc = (B + IRC - 1) / IRC + (2 * B + IRC - 1) / IRC;
    inst_dW_kernel<<<nc, 256, 0, st>>>(B, n_tables, F, inst_grad, O_inst, workspace);
    ER_LAUNCH_CHECK();
    const long long total = 2 * 64LL * F + (long long)(n_tables - 1) * 4096 + 128 + 64 * (n_tables - 1);
    inst_dWred_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(B, n_tables, F, workspace, gscale_dev, o);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_adam_apply_multi(int n_tensors, const elimrec_adam_tensor_t* t, const double* consts_dev,
                                         double beta1, double beta2, float eps, float weight_decay,
                                         elimrec_stream_t stream) {
    ER_CHECK_ARG(n_tensors >= 0 && n_tensors <= ELIMREC_ADAM_MAX_TENSORS, "too many tensors");
    if (n_tensors == 0) return 0;
    AdamMulti a{};
    a.n = n_tensors;
    int blocks = 0;
    for (int i = 0; i < n_tensors; ++i) {
        ER_CHECK_ARG(t[i].row_len > 0 && t[i].grad_ld >= t[i].row_len && t[i].numel >= 0, "bad tensor descriptor");
        a.p[i] = t[i].param; a.g[i] = t[i].grad; a.m[i] = t[i].exp_avg; a.v[i] = t[i].exp_avg_sq;
        a.numel[i] = t[i].numel; a.row_len[i] = t[i].row_len; a.g_ld[i] = t[i].grad_ld;
        a.block_start[i] = blocks;
        blocks += (int)((t[i].numel + ADAM_CHUNK - 1) / ADAM_CHUNK);
    }
    a.block_start[n_tensors] = blocks;
    if (blocks == 0) return 0;
    adam_multi_kernel<<<blocks, 256, 0, er_stream(stream)>>>(a, consts_dev, beta1, beta2, eps, weight_decay);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_inst_forward(int B, int n_tables, int F, const float* O_inst, const float* Wu, const float* Wi,
                                     const float* const* Ws_host, const float* bu, const float* bi, const float* const* bs_host,
                                     float* F_out, float* const* S_out_host, elimrec_stream_t stream) {
    ER_CHECK_ARG(B >= 0 && n_tables >= 1 && n_tables <= 1 + ELIMREC_MAX_MODS && F == 64 * n_tables, "F must be 64 * n_tables");
    ER_CHECK_ARG(O_inst != nullptr && Wu != nullptr && Wi != nullptr && bu != nullptr && bi != nullptr && F_out != nullptr,
                 "NULL operand");
    ER_CHECK_ARG((reinterpret_cast<unsigned long long>(O_inst) & 15) == 0 && (reinterpret_cast<unsigned long long>(Wu) & 15) == 0 &&
                 (reinterpret_cast<unsigned long long>(Wi) & 15) == 0, "operands must be 16-byte aligned");
    if (B == 0) return 0;
    InstFwd w{};
    w.Wu = Wu; w.Wi = Wi; w.bu = bu; w.bi = bi; w.Fout = F_out;
    for (int m = 0; m < n_tables - 1; ++m) {
        ER_CHECK_ARG(Ws_host != nullptr && bs_host != nullptr && S_out_host != nullptr && Ws_host[m] != nullptr &&
                     bs_host[m] != nullptr && S_out_host[m] != nullptr, "NULL head operand");
        ER_CHECK_ARG((reinterpret_cast<unsigned long long>(Ws_host[m]) & 15) == 0, "head weights must be 16-byte aligned");
        w.Ws[m] = Ws_host[m]; w.bs[m] = bs_host[m]; w.Sout[m] = S_out_host[m];
    }
    const int nb = (B + IRB - 1) / IRB + (2 * B + IRB - 1) / IRB;
    inst_fwd_kernel<<<nb, 64 * n_tables, 0, er_stream(stream)>>>(B, n_tables, F, w, O_inst);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_inst_dout_seed(int B, int n_tables, int F, const float* inst_grad, const float* gscale_dev, const float* Wu,
                                     const float* Wi, const float* const* Ws_host, float* dO_inst, const int32_t* rows, int n_mod,
                                     float scale, float* GA, float* GB, int64_t ldg, elimrec_stream_t stream) {
    ER_CHECK_ARG(B >= 0 && n_tables >= 1 && n_tables <= 1 + ELIMREC_MAX_MODS && F == 64 * n_tables, "F must be 64 * n_tables");
    ER_CHECK_ARG(F == 256, "the fused seed epilogue needs one thread per column (F = 256: three modalities)");
    ER_CHECK_ARG(n_mod >= 0 && n_mod <= n_tables - 1 && rows != nullptr && GA != nullptr && GB != nullptr, "bad seed arguments");
    if (B == 0) return 0;
    InstW w{};
    w.Wu = Wu; w.Wi = Wi;
    for (int m = 0; m < n_tables - 1; ++m) w.Ws[m] = Ws_host[m];
    const int nb = (B + IRB - 1) / IRB + (2 * B + IRB - 1) / IRB;
    inst_dO_kernel<<<nb, 256, 0, er_stream(stream)>>>(B, n_tables, F, w, inst_grad, gscale_dev, dO_inst,
                                                       InstSeed{rows, GA, GB, (long long)ldg, scale, n_mod});
    ER_LAUNCH_CHECK();
    return 0;
}
