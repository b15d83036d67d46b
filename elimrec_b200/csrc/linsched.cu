// Kernels of the LINEAR schedule of the training step.
//
// The reference propagates four graphs per step (models/EliMRec.py:250-258): x^g_0 = [E_u ; Z_g] with Z_id = E_i and
// Z_m = X_m W_m^T + b_m for the modality graphs, x^g_k = A_hat^k x^g_0, light_out_g = mean_k x^g_k.  Propagation is linear
// and the item features X_m are constants, so
//
//     light_out_m = mean_k A_hat^k [E_u ; 0]  +  (mean_k A_hat^k [0 ; X_m | 1]) [W_m | b_m]^T
//                 =        (parity part of p)  +           Zbar_m              W'_m^T
//
// with p_k = A_hat^k [E_u ; E_i] (ONE 64-wide propagation, which is also the id graph): A_hat is bipartite, so
// A_hat^k [E_u ; 0] lives on the user rows for even k and on the item rows for odd k, exactly where p_k carries it.
// Zbar_m is computed once; per step the modality blocks cost a [3B x D_m] x [D_m x 64] GEMM on the gathered instance rows
// instead of a feature-streaming projection plus three 256-wide SpMMs.
#include "common.cuh"

namespace {

struct LinLayers {
    int n;                                     // number of tables p_0 .. p_L
    const float* user[ELIMREC_MAX_LAYERS + 1];  // p_k restricted to user rows  [U x 64], row stride user_ld[k]
    const float* item[ELIMREC_MAX_LAYERS + 1];  // p_k restricted to item rows  [I x 64]
    long long user_ld[ELIMREC_MAX_LAYERS + 1];
    long long item_ld[ELIMREC_MAX_LAYERS + 1];
};

// One half-warp per output row: out[j, 0:64] = scale * sum_k p_k[node];  out[j, 64(1+m) : 64(2+m)] (+)= scale * sum over the
// layers whose parity carries the E_u part (users: even k, items: odd k).  Summation order k = 0..L like torch.stack + mean.
__global__ void __launch_bounds__(256)
lin_assemble_kernel(long long n_rows, const int* __restrict__ rows, int num_users, LinLayers lay, float scale, int n_mod,
                    int accumulate, float* __restrict__ out, long long ldo) {
    const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int l = threadIdx.x & 15;
    if (j >= n_rows) return;
    const long long node = (rows != nullptr) ? (long long)__ldg(rows + j) : j;
    const bool is_user = node < num_users;
    const long long r = is_user ? node : node - num_users;
    float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sp = sa;
    const int par = is_user ? 0 : 1;
    for (int k = 0; k < lay.n; ++k) {
        const float* base = is_user ? lay.user[k] + r * lay.user_ld[k] : lay.item[k] + r * lay.item_ld[k];
        const float4 v = __ldg(reinterpret_cast<const float4*>(base) + l);
        if (k == 0) sa = v; else add4(sa, v);
        if ((k & 1) == par) add4(sp, v);
    }
    sa.x *= scale; sa.y *= scale; sa.z *= scale; sa.w *= scale;
    sp.x *= scale; sp.y *= scale; sp.z *= scale; sp.w *= scale;
    float4* o = reinterpret_cast<float4*>(out + j * ldo) + l;
    o[0] = sa;
    for (int m = 0; m < n_mod; ++m) {
        float4 t = sp;
        if (accumulate) add4(t, o[16 * (m + 1)]);
        o[16 * (m + 1)] = t;
    }
}

// Backward of lin_assemble on the instance rows: the gradient that enters layer k of the 64-wide chain,
//   dst[node, 0:64] += scale * ( dO[j, 0:64] + [k even (user) / odd (item)] * sum_{m} dO[j, 64(1+m) : 64(2+m)] ),
// atomically (a node can be sampled several times in one batch).
__global__ void __launch_bounds__(256)
lin_seed_kernel(int n_rows, const int* __restrict__ rows, int num_users, int k, const float* __restrict__ dO, long long ldo,
                int n_mod, float scale, float* __restrict__ dst, long long ldd) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int l = threadIdx.x & 15;
    if (j >= n_rows) return;
    const int node = __ldg(rows + j);
    const int par = (node < num_users) ? 0 : 1;
    const float4* s = reinterpret_cast<const float4*>(dO + (long long)j * ldo) + l;
    float4 v = __ldg(s);
    if ((k & 1) == par)
        for (int m = 0; m < n_mod; ++m) add4(v, __ldg(s + 16 * (m + 1)));
    float* d = dst + (long long)node * ldd + 4 * l;
    atomicAdd(d + 0, scale * v.x);
    atomicAdd(d + 1, scale * v.y);
    atomicAdd(d + 2, scale * v.z);
    atomicAdd(d + 3, scale * v.w);
}

// Both seeds of the backward chain at once: GA[node] += scale * sum of ALL blocks of dO[j], GB[node] += scale * dO[j, 0:64].
// Layer k of the chain then adds GA on the rows whose parity carries the modality graphs' E_u part (users: k even, items:
// k odd) and GB on the others - as the additive epilogue of the propagation launch itself (elimrec_spmm64_pair).
__global__ void __launch_bounds__(256)
lin_seed2_kernel(int n_rows, const int* __restrict__ rows, const float* __restrict__ dO, long long ldo, int n_mod, float scale,
                 float* __restrict__ GA, float* __restrict__ GB, long long ldg) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int l = threadIdx.x & 15;
    if (j >= n_rows) return;
    const long long node = __ldg(rows + j);
    const float4* s = reinterpret_cast<const float4*>(dO + (long long)j * ldo) + l;
    const float4 b = __ldg(s);
    float4 a = b;
    for (int m = 0; m < n_mod; ++m) add4(a, __ldg(s + 16 * (m + 1)));
    float* da = GA + node * ldg + 4 * l;
    float* db = GB + node * ldg + 4 * l;
    atomicAdd(da + 0, scale * a.x); atomicAdd(da + 1, scale * a.y); atomicAdd(da + 2, scale * a.z); atomicAdd(da + 3, scale * a.w);
    atomicAdd(db + 0, scale * b.x); atomicAdd(db + 1, scale * b.y); atomicAdd(db + 2, scale * b.z); atomicAdd(db + 3, scale * b.w);
}

struct PackProj {
    int n;
    const float* W[ELIMREC_MAX_MODS];
    const float* b[ELIMREC_MAX_MODS];
    float* dst[ELIMREC_MAX_MODS];
    int Dm[ELIMREC_MAX_MODS];
    int Kp[ELIMREC_MAX_MODS];
    int start[ELIMREC_MAX_MODS + 1];   // first thread of each tensor
    int round_tf32;
};

// dst_m [64 x Kp] = [ W_m | b_m | 0 ... ] (optionally rounded to TF32, to nearest: the MMA truncates)
__global__ void pack_proj_kernel(const __grid_constant__ PackProj a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.start[a.n]) return;
    int m = 0;
    while (m + 1 < a.n && t >= a.start[m + 1]) ++m;
    const int i = t - a.start[m];
    const int o = i / a.Kp[m], c = i - o * a.Kp[m];
    float v = 0.f;
    if (c < a.Dm[m]) v = __ldg(a.W[m] + (long long)o * a.Dm[m] + c);
    else if (c == a.Dm[m]) v = __ldg(a.b[m] + o);
    if (a.round_tf32) {
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
        v = __uint_as_float(r);
    }
    a.dst[m][i] = v;
}

// Y[r, 0:width] += scale * X[r, 0:width]   (accumulation of the constant Zbar tables at construction)
__global__ void axpy_2d_kernel(long long n_rows, int width4, float scale, const float* __restrict__ X, long long ldx,
                               float* __restrict__ Y, long long ldy, int accumulate) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = t / width4;
    const int c = (int)(t % width4) * 4;
    if (r >= n_rows) return;
    float4 v = __ldg(reinterpret_cast<const float4*>(X + r * ldx + c));
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    float4* y = reinterpret_cast<float4*>(Y + r * ldy + c);
    if (accumulate) add4(v, *y);
    *y = v;
}

// 3xTF32 by concatenation: dst = [hi ; hi ; lo] (pattern 0) or [hi ; lo ; hi] (pattern 1) of src, hi = rna_tf32(x),
// lo = rna_tf32(x - hi).  A TF32 GEMM that reduces over the 3n stacked rows of a pattern-0 and a pattern-1 operand computes
// hi*hi + hi*lo + lo*hi: fp32-class accuracy from the plain TF32 weight-gradient kernel.
__global__ void split3_kernel(long long n_rows, int width4, const float* __restrict__ src, long long lds,
                              float* __restrict__ dst, long long ldd, int pattern) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = t / width4;
    const int c = (int)(t % width4) * 4;
    if (r >= n_rows) return;
    const float4 x = __ldg(reinterpret_cast<const float4*>(src + r * lds + c));
    float4 hi, lo;
    auto sp = [](float v, float& h, float& l) {
        uint32_t a, b;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(v));
        h = __uint_as_float(a);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v - h));
        l = __uint_as_float(b);
    };
    sp(x.x, hi.x, lo.x); sp(x.y, hi.y, lo.y); sp(x.z, hi.z, lo.z); sp(x.w, hi.w, lo.w);
    float4* d0 = reinterpret_cast<float4*>(dst + r * ldd + c);
    float4* d1 = reinterpret_cast<float4*>(dst + (r + n_rows) * ldd + c);
    float4* d2 = reinterpret_cast<float4*>(dst + (r + 2 * n_rows) * ldd + c);
    *d0 = hi;
    *d1 = pattern ? lo : hi;
    *d2 = pattern ? hi : lo;
}

}  // namespace

ELIMREC_API int elimrec_split3_rows(int64_t n_rows, int width, const float* src, int64_t lds, float* dst, int64_t ldd,
                                    int pattern, elimrec_stream_t stream) {
    ER_CHECK_ARG(width % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && (pattern == 0 || pattern == 1), "bad shape");
    if (n_rows <= 0 || width <= 0) return 0;
    const long long threads = n_rows * (width / 4);
    split3_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, er_stream(stream)>>>(n_rows, width / 4, src, lds, dst, ldd, pattern);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_lin_assemble(int64_t n_rows, const int32_t* rows, int32_t num_users, const elimrec_lin_layers_t* layers,
                                     float scale, int n_mod, int accumulate, float* out, int64_t ldo,
                                     elimrec_stream_t stream) {
    ER_CHECK_ARG(layers != nullptr && layers->n >= 1 && layers->n <= ELIMREC_MAX_LAYERS + 1, "1..MAX_LAYERS+1 layer tables");
    ER_CHECK_ARG(n_mod >= 0 && n_mod <= ELIMREC_MAX_MODS && ldo % 4 == 0 && ldo >= 64 * (1 + n_mod), "bad output shape");
    if (n_rows <= 0) return 0;
    LinLayers lay{};
    lay.n = layers->n;
    for (int k = 0; k < lay.n; ++k) {
        ER_CHECK_ARG(layers->user_ld[k] % 4 == 0 && layers->item_ld[k] % 4 == 0, "row strides must be multiples of 4 floats");
        lay.user[k] = layers->user[k]; lay.item[k] = layers->item[k];
        lay.user_ld[k] = layers->user_ld[k]; lay.item_ld[k] = layers->item_ld[k];
    }
    const long long threads = n_rows * 16;
    lin_assemble_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, er_stream(stream)>>>(n_rows, rows, num_users, lay, scale,
                                                                                          n_mod, accumulate, out, ldo);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_lin_seed(int n_rows, const int32_t* rows, int32_t num_users, int layer, const float* dO, int64_t ldo,
                                 int n_mod, float scale, float* dst, int64_t ldd, elimrec_stream_t stream) {
    ER_CHECK_ARG(rows != nullptr && dO != nullptr && dst != nullptr, "NULL buffer");
    ER_CHECK_ARG(n_mod >= 0 && n_mod <= ELIMREC_MAX_MODS && ldo % 4 == 0 && ldd % 4 == 0, "bad shape");
    if (n_rows <= 0) return 0;
    lin_seed_kernel<<<(n_rows * 16 + 255) / 256, 256, 0, er_stream(stream)>>>(n_rows, rows, num_users, layer, dO, ldo, n_mod,
                                                                             scale, dst, ldd);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_lin_seed2(int n_rows, const int32_t* rows, const float* dO, int64_t ldo, int n_mod, float scale,
                                  float* GA, float* GB, int64_t ldg, elimrec_stream_t stream) {
    ER_CHECK_ARG(rows != nullptr && dO != nullptr && GA != nullptr && GB != nullptr, "NULL buffer");
    ER_CHECK_ARG(n_mod >= 0 && n_mod <= ELIMREC_MAX_MODS && ldo % 4 == 0 && ldg % 4 == 0, "bad shape");
    if (n_rows <= 0) return 0;
    lin_seed2_kernel<<<(n_rows * 16 + 255) / 256, 256, 0, er_stream(stream)>>>(n_rows, rows, dO, ldo, n_mod, scale, GA, GB, ldg);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_pack_proj_weights(int n, const elimrec_pack_proj_t* t, int round_tf32, elimrec_stream_t stream) {
    ER_CHECK_ARG(n >= 0 && n <= ELIMREC_MAX_MODS && (n == 0 || t != nullptr), "at most MAX_MODS projections");
    if (n == 0) return 0;
    PackProj a{};
    a.n = n;
    a.round_tf32 = round_tf32;
    int total = 0;
    for (int m = 0; m < n; ++m) {
        ER_CHECK_ARG(t[m].Kp > t[m].Dm && t[m].Dm > 0, "padded width must exceed the feature width (bias column)");
        a.W[m] = t[m].W; a.b[m] = t[m].b; a.dst[m] = t[m].dst; a.Dm[m] = (int)t[m].Dm; a.Kp[m] = (int)t[m].Kp;
        a.start[m] = total;
        total += 64 * (int)t[m].Kp;
    }
    a.start[n] = total;
    pack_proj_kernel<<<(total + 255) / 256, 256, 0, er_stream(stream)>>>(a);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_axpy_2d(int64_t n_rows, int width, float scale, const float* X, int64_t ldx, float* Y, int64_t ldy,
                                int accumulate, elimrec_stream_t stream) {
    ER_CHECK_ARG(width % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "width/strides must be multiples of 4");
    if (n_rows <= 0 || width <= 0) return 0;
    const long long threads = n_rows * (width / 4);
    axpy_2d_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, er_stream(stream)>>>(n_rows, width / 4, scale, X, ldx, Y, ldy,
                                                                                     accumulate);
    ER_LAUNCH_CHECK();
    return 0;
}
