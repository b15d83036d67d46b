// Column-block helpers of the non-default model variants.  All HBM-bound
// elementwise passes: one thread per float4, fully coalesced, grid sized from the element count.
//   tie_blocks   mm_fusion_mode='mean' (reference models/EliMRec.py:224-225): W.mean(stack(reps)) is evaluated as the
//                concat contraction with the tied weight [W/G | W/G | ...]  (G a power of two: exact scaling)
//   fold_blocks  the transpose of tie_blocks (gradient of the tied weight; d E_u = sum over the graphs' blocks)
//   axpy_rows    the self-loop term of adj_type 'norm' / 'mean' (models/EliMRec.py:332-352): Y[r,:] += s[r] * X[r,:]
//   layer_mean   torch.mean(torch.stack(embs, 1), 1) (models/EliMRec.py:246-247) when it cannot be fused into the
//                last SpMM (self loops)
#include "common.cuh"

namespace {

__global__ void tie_blocks_kernel(long long n_rows, const float* __restrict__ src, long long src_ld,
                                  float* __restrict__ dst, long long dst_ld, int n_rep, float scale) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = t >> 4;
    const int c = (int)(t & 15) * 4;
    if (r >= n_rows) return;
    float4 v = __ldg(reinterpret_cast<const float4*>(src + r * src_ld + c));
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    for (int g = 0; g < n_rep; ++g) *reinterpret_cast<float4*>(dst + r * dst_ld + g * 64 + c) = v;
}

__global__ void fold_blocks_kernel(long long n_rows, const float* __restrict__ src, long long src_ld, int n_rep,
                                   float scale, float* __restrict__ dst, long long dst_ld) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = t >> 4;
    const int c = (int)(t & 15) * 4;
    if (r >= n_rows) return;
    float4 a = __ldg(reinterpret_cast<const float4*>(src + r * src_ld + c));
    for (int g = 1; g < n_rep; ++g) add4(a, __ldg(reinterpret_cast<const float4*>(src + r * src_ld + g * 64 + c)));
    a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
    *reinterpret_cast<float4*>(dst + r * dst_ld + c) = a;
}

__global__ void axpy_rows_kernel(long long n_rows, int width4, const float* __restrict__ s, const float* __restrict__ X,
                                 long long ldx, float* __restrict__ Y, long long ldy) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = t / width4;
    const int c = (int)(t % width4) * 4;
    if (r >= n_rows) return;
    const float w = __ldg(s + r);
    const float4 x = __ldg(reinterpret_cast<const float4*>(X + r * ldx + c));
    float4 y = *reinterpret_cast<float4*>(Y + r * ldy + c);
    fma4(y, w, x);
    *reinterpret_cast<float4*>(Y + r * ldy + c) = y;
}

struct LayerPtrs {
    const float* x[ELIMREC_MAX_LAYERS + 1];
    long long ld[ELIMREC_MAX_LAYERS + 1];
};

__global__ void layer_mean_kernel(long long n_rows, int width4, int n_layers, LayerPtrs p, float scale,
                                  float* __restrict__ out, long long ld_out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = t / width4;
    const int c = (int)(t % width4) * 4;
    if (r >= n_rows) return;
    float4 a = __ldg(reinterpret_cast<const float4*>(p.x[0] + r * p.ld[0] + c));
    for (int k = 1; k < n_layers; ++k) add4(a, __ldg(reinterpret_cast<const float4*>(p.x[k] + r * p.ld[k] + c)));
    a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
    *reinterpret_cast<float4*>(out + r * ld_out + c) = a;
}

inline unsigned blocks_for(long long threads) { return (unsigned)((threads + 255) / 256); }

}  // namespace

ELIMREC_API int elimrec_tie_blocks(int64_t n_rows, const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int n_rep,
                                   float scale, elimrec_stream_t stream) {
    ER_CHECK_ARG(src_ld % 4 == 0 && dst_ld % 4 == 0, "strides must be multiples of 4");
    ER_CHECK_ARG(n_rep >= 1 && dst_ld >= 64LL * n_rep, "dst row too short for n_rep blocks of 64");
    if (n_rows <= 0) return 0;
    tie_blocks_kernel<<<blocks_for(n_rows * 16), 256, 0, er_stream(stream)>>>(n_rows, src, src_ld, dst, dst_ld, n_rep, scale);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_fold_blocks(int64_t n_rows, const float* src, int64_t src_ld, int n_rep, float scale, float* dst,
                                    int64_t dst_ld, elimrec_stream_t stream) {
    ER_CHECK_ARG(src_ld % 4 == 0 && dst_ld % 4 == 0, "strides must be multiples of 4");
    ER_CHECK_ARG(n_rep >= 1 && src_ld >= 64LL * n_rep, "src row too short for n_rep blocks of 64");
    if (n_rows <= 0) return 0;
    fold_blocks_kernel<<<blocks_for(n_rows * 16), 256, 0, er_stream(stream)>>>(n_rows, src, src_ld, n_rep, scale, dst, dst_ld);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_axpy_rows(int64_t n_rows, int width, const float* row_scale, const float* X, int64_t ldx, float* Y,
                                  int64_t ldy, elimrec_stream_t stream) {
    ER_CHECK_ARG(width > 0 && width % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "width/strides must be multiples of 4");
    ER_CHECK_ARG(row_scale != nullptr, "row_scale is required");
    if (n_rows <= 0) return 0;
    axpy_rows_kernel<<<blocks_for(n_rows * (width / 4)), 256, 0, er_stream(stream)>>>(n_rows, width / 4, row_scale, X, ldx, Y,
                                                                                    ldy);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_layer_mean(int64_t n_rows, int width, int n_layers, const float* const* layers_host,
                                   const int64_t* ld_host, float scale, float* out, int64_t ld_out, elimrec_stream_t stream) {
    ER_CHECK_ARG(width > 0 && width % 4 == 0 && ld_out % 4 == 0, "width/strides must be multiples of 4");
    ER_CHECK_ARG(n_layers >= 1 && n_layers <= ELIMREC_MAX_LAYERS + 1, "n_layers out of range");
    if (n_rows <= 0) return 0;
    LayerPtrs p{};
    for (int k = 0; k < n_layers; ++k) {
        ER_CHECK_ARG(layers_host[k] != nullptr && ld_host[k] % 4 == 0, "bad layer pointer / stride");
        p.x[k] = layers_host[k];
        p.ld[k] = ld_host[k];
    }
    layer_mean_kernel<<<blocks_for(n_rows * (width / 4)), 256, 0, er_stream(stream)>>>(n_rows, width / 4, n_layers, p, scale,
                                                                                     out, ld_out);
    ER_LAUNCH_CHECK();
    return 0;
}
