// Graph propagation: Y = A_hat * X over a CSR half of the normalised bipartite adjacency.
// Replaces torch.sparse.mm(norm_adj, all_emb) (reference models/EliMRec.py:244) and its backward.
//
// Mapping: one warp per SEGMENT (<= seg_len consecutive edges of one row).  A lane owns 4 (F=64,
// two edges per warp-iteration on the two half-warps), 4 (F=128) or 8 (F=256) consecutive-by-128
// output columns, so every neighbour row is fetched as fully coalesced 16-byte vector loads
// (256 B / 512 B / 1 KB per row).  UNR independent row fetches are kept in flight per lane to cover
// L2 latency - the operand slab (48-115 MB at the single-GPU shapes) is L2 resident.
// Rows longer than seg_len are split over whole CTAs (8 segments each, reduced in shared memory); the
// last-arriving CTA of a row adds the per-CTA partial sums in a FIXED order, so results are
// bit-reproducible run to run (no float atomics).
#include "common.cuh"

namespace {

struct MeanEpi {
    int n_prev;
    const float* prev[ELIMREC_MAX_LAYERS];
    long long prev_ld[ELIMREC_MAX_LAYERS];
    int prev_width[ELIMREC_MAX_LAYERS];
    float* out;
    long long ld;
    int width;
    float scale;
};

// optional additive epilogue of the masked variants: Y[row] += g[row] where mask[row] != 0 (mask NULL: every row).
// Backward pass: the row-sparse layer-mean gradient G enters every layer of the chain (d x_{k-1} = A^T d x_k + G); G lives
// in a slab that is only valid on the <= 3B instance rows, which is what `mask` marks.
struct AddEpi {
    const float* g;
    long long ld;
    const unsigned char* mask;
};

template <int F>
__device__ __forceinline__ void seg_add(int row, float4 (&acc)[(F >= 128) ? F / 128 : 1], const AddEpi& add, int lane) {
    constexpr int NV = (F >= 128) ? F / 128 : 1;
    if (add.g == nullptr || (add.mask != nullptr && __ldg(add.mask + row) == 0)) return;
    const int l = (F == 64) ? (lane & 15) : lane;
    const float4* gp = reinterpret_cast<const float4*>(add.g + (long long)row * add.ld) + l;
#pragma unroll
    for (int i = 0; i < NV; ++i) add4(acc[i], __ldg(gp + i * 32));
}

// HEAVY = false: segments [seg_base, n_seg) are whole rows (lean path, tuned for FULL occupancy: 32 registers, 64 warps/SM -
// the gather is latency-bound until ~11 TB/s of L2->SM traffic, measured: 2 fetches in flight x 64 warps beats 8 x 16).
// HEAVY = true : segments [0, n_heavy_seg) belong to split rows, one CTA = 8 segments of one row.
// accumulate one segment: acc += sum_e val[e] * X[col[e], :]   (col_mask != NULL: edges to unmarked columns are dropped
// BEFORE the gather by an order-preserving compaction onto lanes 0..cnt-1)
template <int F, int UNR, bool MASK>
__device__ __forceinline__ void seg_accumulate(const int4 sg, const int* __restrict__ col, const float* __restrict__ val,
                                               const float* __restrict__ X, long long ldx,
                                               const unsigned char* __restrict__ col_mask, int lane,
                                               float4 (&acc)[(F >= 128) ? F / 128 : 1]) {
    constexpr int EPW = (F == 64) ? 2 : 1;          // edges per warp-iteration
    constexpr int NV = (F >= 128) ? F / 128 : 1;    // float4 per lane
    const unsigned full = 0xffffffffu;
    const int sub = (F == 64) ? (lane >> 4) : 0;
    const int l = (F == 64) ? (lane & 15) : lane;
    for (int base = sg.y; base < sg.z; base += 32) {
        const int e = base + lane;
        int c = 0;
        float w = 0.f;
        if (e < sg.z) {
            c = __ldg(col + e);
            w = __ldg(val + e);
        }
        int cnt = min(32, sg.z - base);
        if (MASK && col_mask != nullptr) {
            const unsigned alive = __ballot_sync(full, e < sg.z && __ldg(col_mask + c) != 0);
            cnt = __popc(alive);
            if (cnt == 0) continue;
            const int src = __fns(alive, 0, lane + 1) & 31;
            c = __shfl_sync(full, c, src);
            w = __shfl_sync(full, w, src);
        }
        for (int j = 0; j < cnt; j += UNR * EPW) {
            float4 v[UNR][NV];
            float ww[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int jj = j + u * EPW + sub;
                const int cc = __shfl_sync(full, c, jj & 31);
                const float wv = __shfl_sync(full, w, jj & 31);
                const bool ok = jj < cnt;
                ww[u] = ok ? wv : 0.f;
                const float4* p = reinterpret_cast<const float4*>(X + (long long)cc * ldx) + l;
#pragma unroll
                for (int i = 0; i < NV; ++i) v[u][i] = ok ? __ldg(p + i * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u)
#pragma unroll
                for (int i = 0; i < NV; ++i) fma4(acc[i], ww[u], v[u][i]);
        }
    }
    if (F == 64) {  // fold the two half-warps; afterwards both halves hold the full row
        acc[0].x += __shfl_xor_sync(full, acc[0].x, 16);
        acc[0].y += __shfl_xor_sync(full, acc[0].y, 16);
        acc[0].z += __shfl_xor_sync(full, acc[0].z, 16);
        acc[0].w += __shfl_xor_sync(full, acc[0].w, 16);
    }
}

// write one finished row: Y[row] = acc and / or the fused layer-mean epilogue
template <int F, bool MEAN>
__device__ __forceinline__ void seg_store(int row, const float4 (&acc)[(F >= 128) ? F / 128 : 1], float* __restrict__ Y,
                                          long long ldy, const MeanEpi& epi, int lane) {
    constexpr int NV = (F >= 128) ? F / 128 : 1;
    const int sub = (F == 64) ? (lane >> 4) : 0;
    const int l = (F == 64) ? (lane & 15) : lane;
    if (Y != nullptr && (F != 64 || lane < 16)) {
        float4* yp = reinterpret_cast<float4*>(Y + (long long)row * ldy) + l;
#pragma unroll
        for (int i = 0; i < NV; ++i) yp[i * 32] = acc[i];
    }
    if (MEAN) {
        // fused torch.mean(torch.stack(layers, 1), 1): ((x0 + x1) + ...) + x_L, then * 1/(L+1)
        if (F == 64) {
            const int c = l * 4;
            for (int g = sub; g * 64 < epi.width; g += 2) {
                const int oc = g * 64 + c;
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int k = 0; k < epi.n_prev; ++k) {
                    const int pc = (epi.prev_width[k] == 64) ? c : oc;
                    const float4 t = __ldg(reinterpret_cast<const float4*>(epi.prev[k] + (long long)row * epi.prev_ld[k] + pc));
                    if (k == 0) s = t; else add4(s, t);
                }
                add4(s, acc[0]);
                s.x *= epi.scale; s.y *= epi.scale; s.z *= epi.scale; s.w *= epi.scale;
                *reinterpret_cast<float4*>(epi.out + (long long)row * epi.ld + oc) = s;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int oc = (l + i * 32) * 4;
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int k = 0; k < epi.n_prev; ++k) {
                    const int pc = (epi.prev_width[k] == 64) ? (oc & 63) : oc;
                    const float4 t = __ldg(reinterpret_cast<const float4*>(epi.prev[k] + (long long)row * epi.prev_ld[k] + pc));
                    if (k == 0) s = t; else add4(s, t);
                }
                add4(s, acc[i]);
                s.x *= epi.scale; s.y *= epi.scale; s.z *= epi.scale; s.w *= epi.scale;
                *reinterpret_cast<float4*>(epi.out + (long long)row * epi.ld + oc) = s;
            }
        }
    }
}

// MASK = true (row-sparse last layers of a training step, see elimrec_spmm_masked):
//   row_mask != NULL : segments of rows with row_mask[row] == 0 are skipped, their output is left untouched
//   col_mask != NULL : edges whose column has col_mask[col] == 0 are dropped BEFORE the gather (their X rows are
//                      never read); the surviving edges keep their order, so the sums equal the unmasked ones
//                      whenever the dropped rows of X are zero.
template <int F, bool MEAN, bool HEAVY, int UNR_, int MINB, bool MASK = false>
__global__ void __launch_bounds__(256, MINB)
spmm_seg_kernel(int seg_base, int n_seg, const int4* __restrict__ seg, const int2* __restrict__ heavy, int* __restrict__ counter,
                const int* __restrict__ col, const float* __restrict__ val, const float* __restrict__ X,
                long long ldx, float* __restrict__ Y, long long ldy, float* __restrict__ partial, MeanEpi epi,
                const unsigned char* __restrict__ row_mask = nullptr, const unsigned char* __restrict__ col_mask = nullptr,
                AddEpi add = AddEpi{nullptr, 0, nullptr}) {
    constexpr int NV = (F >= 128) ? F / 128 : 1;    // float4 per lane
    const unsigned full = 0xffffffffu;
    const int warp = seg_base + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (warp >= n_seg) return;
    const int4 sg = __ldg(seg + warp);
    if (MASK && row_mask != nullptr && __ldg(row_mask + sg.x) == 0) return;   // HEAVY: the 8 warps of a CTA share the row
    const int l = (F == 64) ? (lane & 15) : lane;

    float4 acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    seg_accumulate<F, UNR_, MASK>(sg, col, val, X, ldx, col_mask, lane, acc);

    if (HEAVY) {
        // Split row.  Its segment count is padded to a multiple of 8 on the host, so all 8 warps of this CTA work on
        // the SAME row: reduce them in shared memory first (fixed order), then one partial per CTA goes to scratch and
        // the last-arriving CTA of the row (atomic counter) adds the CTA partials in CTA order.  Deterministic.
        __shared__ float4 red[8][F / 4];
        const int wib = threadIdx.x >> 5;
        if (F != 64 || lane < 16) {
#pragma unroll
            for (int i = 0; i < NV; ++i) red[wib][l + i * 32] = acc[i];
        }
        __syncthreads();
        if (wib != 0) return;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float4 t = red[0][l + i * 32];
#pragma unroll
            for (int w2 = 1; w2 < 8; ++w2) add4(t, red[w2][l + i * 32]);
            acc[i] = t;
        }
        const int2 hv = __ldg(heavy + sg.w);  // {first CTA of the row, number of CTAs}
        if (hv.y > 1) {
            float4* pp = reinterpret_cast<float4*>(partial + (long long)blockIdx.x * F);
            if (F != 64 || lane < 16) {
#pragma unroll
                for (int i = 0; i < NV; ++i) pp[l + i * 32] = acc[i];
            }
            __threadfence();
            int old = 0;
            if (lane == 0) old = atomicAdd(counter + sg.w, 1);
            old = __shfl_sync(full, old, 0);
            if (old != hv.y - 1) return;
            __threadfence();
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4* ps = reinterpret_cast<const float4*>(partial + (long long)hv.x * F) + l;
            constexpr int RU = 4;
            for (int s2 = 0; s2 < hv.y; s2 += RU) {
                float4 t[RU][NV];
#pragma unroll
                for (int u = 0; u < RU; ++u)
#pragma unroll
                    for (int i = 0; i < NV; ++i)
                        t[u][i] = (s2 + u < hv.y) ? __ldcg(ps + (long long)(s2 + u) * (F / 4) + i * 32)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < RU; ++u)
#pragma unroll
                    for (int i = 0; i < NV; ++i) add4(acc[i], t[u][i]);
            }
            if (lane == 0) counter[sg.w] = 0;  // leave the counter ready for the next launch
        }
    }

    if (MASK) seg_add<F>(sg.x, acc, add, lane);
    seg_store<F, MEAN>(sg.x, acc, Y, ldy, epi, lane);
}

// Whole rows under a SPARSE mask (5-50 % of the rows marked): a CTA owns SEGS consecutive segments; its threads fetch the
// segment descriptors and mask bytes together (two dependent loads per SEGS segments instead of per segment), the marked
// ones are compacted into shared memory and dealt round-robin to the 8 warps.  The grid is SEGS/8 times smaller than
// one-warp-per-segment, fits in one or two waves, and so really overlaps with the other launches of the layer; 4 row
// fetches in flight per lane because few warps are active (8 cost occupancy: measured slower).  (The list order varies run to run, the results do not: every
// row is still produced by exactly one warp, edges in CSR order.)
template <int F, bool MEAN, int SEGS>
__global__ void __launch_bounds__(256, 2)
spmm_light_sparse_kernel(int seg_base, int n_seg, const int4* __restrict__ seg, const int* __restrict__ col,
                         const float* __restrict__ val, const float* __restrict__ X, long long ldx, float* __restrict__ Y,
                         long long ldy, MeanEpi epi, const unsigned char* __restrict__ row_mask,
                         const unsigned char* __restrict__ col_mask, AddEpi add) {
    constexpr int NV = (F >= 128) ? F / 128 : 1;
    __shared__ int4 list[SEGS];
    __shared__ int n_list;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int first = seg_base + blockIdx.x * SEGS;
    if (threadIdx.x == 0) n_list = 0;
    __syncthreads();
    if (threadIdx.x < SEGS && first + threadIdx.x < n_seg) {
        const int4 mine = __ldg(seg + first + threadIdx.x);
        if (row_mask == nullptr || __ldg(row_mask + mine.x) != 0) list[atomicAdd(&n_list, 1)] = mine;
    }
    __syncthreads();
    const int n = n_list;
    for (int i = wib; i < n; i += 8) {
        const int4 sg = list[i];
        float4 acc[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        seg_accumulate<F, 4, true>(sg, col, val, X, ldx, col_mask, lane, acc);
        seg_add<F>(sg.x, acc, add, lane);
        seg_store<F, MEAN>(sg.x, acc, Y, ldy, epi, lane);
    }
}

__global__ void scatter_add_rows_kernel(int n_rows, const int* __restrict__ rows, int row_lo, int row_hi, int row_off,
                                        const float* __restrict__ src, long long src_ld, int fold, float* dst,
                                        long long dst_ld, int width, float scale) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    const int node = __ldg(rows + r);
    if (node < row_lo || node >= row_hi) return;
    const float* s = src + (long long)r * src_ld;
    float* d = dst + (long long)(node - row_off) * dst_ld;
    for (int c = lane; c < width; c += 32) {
        float v = 0.f;
        for (int f = 0; f < fold; ++f) v += s[f * width + c];
        atomicAdd(d + c, scale * v);
    }
}

__global__ void mark_rows_kernel(int n_rows, const int* __restrict__ rows, unsigned char* __restrict__ mask) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rows) mask[__ldg(rows + r)] = 1;
}

// rows = [users | U + pos | U + neg] (the instance rows of a BPR batch) and mask[rows] = 1, one launch
__global__ void inst_rows_kernel(int B, const long long* __restrict__ users, const long long* __restrict__ pos,
                                 const long long* __restrict__ neg, int num_users, int* __restrict__ rows,
                                 unsigned char* __restrict__ mask, unsigned char* __restrict__ mask2) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= 3 * B) return;
    const int k = r / B, b = r - k * B;
    const int node = (k == 0) ? (int)__ldg(users + b) : num_users + (int)__ldg((k == 1 ? pos : neg) + b);
    rows[r] = node;
    if (mask != nullptr) mask[node] = 1;
    if (mask2 != nullptr) mask2[node] = 1;
}

// out_mask[col] = 1 for every edge (row, col) of a marked row: the rows of the other side that the marked rows read
__global__ void mark_neighbors_kernel(int n_seg, const int4* __restrict__ seg, const int* __restrict__ col,
                                      const unsigned char* __restrict__ row_mask, unsigned char* __restrict__ out_mask) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_seg) return;
    const int4 sg = __ldg(seg + w);
    if (__ldg(row_mask + sg.x) == 0) return;
    for (int e = sg.y + lane; e < sg.z; e += 32) out_mask[__ldg(col + e)] = 1;
}

__global__ void zero_rows_kernel(int n_rows, const int* __restrict__ rows, int row_lo, int row_hi, int row_off,
                                 float* __restrict__ dst, long long dst_ld, int width4) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    const int node = __ldg(rows + r);
    if (node < row_lo || node >= row_hi) return;
    float4* d = reinterpret_cast<float4*>(dst + (long long)(node - row_off) * dst_ld);
    for (int c = lane; c < width4; c += 32) d[c] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void gather_rows_kernel(int n_rows, const int* __restrict__ rows, const float* __restrict__ src,
                                   long long src_ld, float* __restrict__ dst, long long dst_ld, int width4) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    const float4* s = reinterpret_cast<const float4*>(src + (long long)__ldg(rows + r) * src_ld);
    float4* d = reinterpret_cast<float4*>(dst + (long long)r * dst_ld);
    for (int c = lane; c < width4; c += 32) d[c] = __ldg(s + c);
}

__global__ void broadcast_cols_kernel(long long n_rows, const float* __restrict__ src, long long src_ld,
                                      float* __restrict__ dst, long long dst_ld, int n_rep) {
    // one thread per (row, float4 of the 64 source columns)
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = t >> 4;
    const int c = (int)(t & 15) * 4;
    if (r >= n_rows) return;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + r * src_ld + c));
    for (int g = 0; g < n_rep; ++g) *reinterpret_cast<float4*>(dst + r * dst_ld + g * 64 + c) = v;
}

__global__ void copy_2d_kernel(long long n_rows, int width4, const float* __restrict__ src, long long src_ld,
                               float* __restrict__ dst, long long dst_ld) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = t / width4;
    const int c = (int)(t % width4) * 4;
    if (r >= n_rows) return;
    *reinterpret_cast<float4*>(dst + r * dst_ld + c) = __ldg(reinterpret_cast<const float4*>(src + r * src_ld + c));
}

}  // namespace

namespace {
int spmm_launch(int width, int part, int n_seg, int n_heavy_seg, const int32_t* seg, const int32_t* heavy, int32_t* counter,
                const int32_t* col, const float* val, const float* X, int64_t ldx, float* Y, int64_t ldy,
                float* partial, const elimrec_mean_epilogue_t* epi, const unsigned char* row_mask,
                const unsigned char* col_mask, int density_hint, elimrec_stream_t stream, const float* addend = nullptr,
                long long ld_add = 0, const unsigned char* add_mask = nullptr) {
    // whole rows under a mask: segments per CTA - 256 when only a few % of the rows are marked, else 32
    const int chunk = (row_mask != nullptr && density_hint <= 10) ? 256 : 32;
    ER_CHECK_ARG(width == 64 || width == 128 || width == 256, "width must be 64, 128 or 256");
    ER_CHECK_ARG(ldx % 4 == 0 && (Y == nullptr || ldy % 4 == 0), "row strides must be multiples of 4 floats");
    ER_CHECK_ARG(Y != nullptr || (epi != nullptr && epi->mean_out != nullptr), "no output requested");
    if (n_seg <= 0) return 0;
    MeanEpi me{};
    if (epi != nullptr && epi->mean_out != nullptr) {
        ER_CHECK_ARG(epi->n_prev >= 0 && epi->n_prev <= ELIMREC_MAX_LAYERS, "too many layers");
        ER_CHECK_ARG(epi->mean_width == width || width == 64, "mean epilogue: width mismatch");
        ER_CHECK_ARG(epi->mean_width % 64 == 0 && epi->mean_ld % 4 == 0, "mean epilogue: bad output shape");
        me.n_prev = epi->n_prev;
        for (int k = 0; k < epi->n_prev; ++k) {
            ER_CHECK_ARG(epi->prev_width[k] == 64 || epi->prev_width[k] == epi->mean_width, "mean epilogue: layer width");
            me.prev[k] = epi->prev[k];
            me.prev_ld[k] = epi->prev_ld[k];
            me.prev_width[k] = epi->prev_width[k];
        }
        me.out = epi->mean_out;
        me.ld = epi->mean_ld;
        me.width = epi->mean_width;
        me.scale = epi->mean_scale;
    }
    const int4* sg = reinterpret_cast<const int4*>(seg);
    const int2* hv = reinterpret_cast<const int2*>(heavy);
    cudaStream_t st = er_stream(stream);
    const bool mean = me.out != nullptr;
    const bool masked = row_mask != nullptr || col_mask != nullptr || addend != nullptr;
    ER_CHECK_ARG(addend == nullptr || ld_add % 4 == 0, "addend stride must be a multiple of 4 floats");
    const AddEpi add{addend, ld_add, add_mask};
    ER_CHECK_ARG(n_heavy_seg >= 0 && n_heavy_seg % 8 == 0 && n_heavy_seg <= n_seg, "n_heavy_seg must be a multiple of 8");
    const int hb = n_heavy_seg / 8;                      // CTAs of split rows
    const int lb = (n_seg - n_heavy_seg + 7) / 8;        // CTAs of whole rows
#define LAUNCH(F, MEAN, MASK)                                                                                             \
    do {                                                                                                                  \
        if (hb > 0 && part != 2)                                                                                          \
            spmm_seg_kernel<F, MEAN, true, 2, 5, MASK><<<hb, 256, 0, st>>>(0, n_heavy_seg, sg, hv, counter, col, val, X,  \
                                                                           ldx, Y, ldy, partial, me, row_mask, col_mask, add); \
        if (lb > 0 && part != 1) {                                                                                        \
            /* measured: narrow col-mask-only is faster one warp per segment; so is anything without a mask to compact */ \
            if (!MASK || (row_mask == nullptr && (F == 64 || col_mask == nullptr)))                                       \
                spmm_seg_kernel<F, MEAN, false, 2, (MEAN ? 6 : 8), MASK><<<lb, 256, 0, st>>>(                             \
                    n_heavy_seg, n_seg, sg, hv, counter, col, val, X, ldx, Y, ldy, partial, me, nullptr, col_mask, add);  \
            else if (chunk == 256)                                                                                        \
                spmm_light_sparse_kernel<F, MEAN, 256><<<(n_seg - n_heavy_seg + 255) / 256, 256, 0, st>>>(                \
                    n_heavy_seg, n_seg, sg, col, val, X, ldx, Y, ldy, me, row_mask, col_mask, add);                       \
            else                                                                                                          \
                spmm_light_sparse_kernel<F, MEAN, 32><<<(n_seg - n_heavy_seg + 31) / 32, 256, 0, st>>>(                   \
                    n_heavy_seg, n_seg, sg, col, val, X, ldx, Y, ldy, me, row_mask, col_mask, add);                       \
        }                                                                                                                 \
    } while (0)
#define LAUNCH_W(F)                                                                  \
    do {                                                                             \
        if (masked) { if (mean) LAUNCH(F, true, true); else LAUNCH(F, false, true); } \
        else { if (mean) LAUNCH(F, true, false); else LAUNCH(F, false, false); }      \
    } while (0)
    if (width == 64) LAUNCH_W(64);
    else if (width == 128) LAUNCH_W(128);
    else LAUNCH_W(256);
#undef LAUNCH_W
#undef LAUNCH
    ER_LAUNCH_CHECK();
    return 0;
}
}  // namespace

ELIMREC_API int elimrec_spmm(int width, int part, int n_seg, int n_heavy_seg, const int32_t* seg, const int32_t* heavy, int32_t* counter,
                             const int32_t* col, const float* val, const float* X, int64_t ldx, float* Y, int64_t ldy,
                             float* partial, const elimrec_mean_epilogue_t* epi, elimrec_stream_t stream) {
    return spmm_launch(width, part, n_seg, n_heavy_seg, seg, heavy, counter, col, val, X, ldx, Y, ldy, partial, epi, nullptr,
                       nullptr, 100, stream);
}

ELIMREC_API int elimrec_spmm_masked(int width, int part, int n_seg, int n_heavy_seg, const int32_t* seg, const int32_t* heavy,
                                    int32_t* counter, const int32_t* col, const float* val, const float* X, int64_t ldx,
                                    float* Y, int64_t ldy, float* partial, const elimrec_mean_epilogue_t* epi,
                                    const uint8_t* row_mask, const uint8_t* col_mask, int row_density_pct,
                                    const float* addend, int64_t ld_add, const uint8_t* add_mask, elimrec_stream_t stream) {
    return spmm_launch(width, part, n_seg, n_heavy_seg, seg, heavy, counter, col, val, X, ldx, Y, ldy, partial, epi, row_mask,
                       col_mask, row_density_pct, stream, addend, ld_add, add_mask);
}

ELIMREC_API int elimrec_mark_rows(int n_rows, const int32_t* rows, int64_t n_nodes, uint8_t* mask, elimrec_stream_t stream) {
    ER_CHECK_ARG(n_nodes >= 0 && mask != nullptr, "mask required");
    cudaStream_t st = er_stream(stream);
    if (n_nodes > 0 && cudaMemsetAsync(mask, 0, (size_t)n_nodes, st) != cudaSuccess) {
        elimrec_set_error("elimrec_mark_rows: memset failed");
        return -3;
    }
    if (n_rows <= 0) return 0;
    mark_rows_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(n_rows, rows, mask);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_inst_rows(int B, const int64_t* users, const int64_t* pos, const int64_t* neg, int32_t num_users,
                                  int32_t* rows, int64_t n_nodes, uint8_t* mask, uint8_t* mask2, elimrec_stream_t stream) {
    ER_CHECK_ARG(B >= 0 && rows != nullptr, "rows required");
    cudaStream_t st = er_stream(stream);
    for (uint8_t* m : {mask, mask2}) {
        if (m != nullptr && n_nodes > 0 && cudaMemsetAsync(m, 0, (size_t)n_nodes, st) != cudaSuccess) {
            elimrec_set_error("elimrec_inst_rows: memset failed");
            return -3;
        }
    }
    if (B == 0) return 0;
    inst_rows_kernel<<<(3 * B + 255) / 256, 256, 0, st>>>(B, (const long long*)users, (const long long*)pos,
                                                         (const long long*)neg, num_users, rows, mask, mask2);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_mark_neighbors(int n_seg, const int32_t* seg, const int32_t* col, const uint8_t* row_mask,
                                       uint8_t* out_mask, elimrec_stream_t stream) {
    ER_CHECK_ARG(row_mask != nullptr && out_mask != nullptr, "masks required");
    if (n_seg <= 0) return 0;
    mark_neighbors_kernel<<<(n_seg + 7) / 8, 256, 0, er_stream(stream)>>>(n_seg, reinterpret_cast<const int4*>(seg), col,
                                                                         row_mask, out_mask);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_zero_rows(int n_rows, const int32_t* rows, int32_t row_lo, int32_t row_hi, int32_t row_offset,
                                  float* dst, int64_t dst_ld, int width, elimrec_stream_t stream) {
    ER_CHECK_ARG(width % 4 == 0 && dst_ld % 4 == 0, "width/stride must be multiples of 4");
    if (n_rows <= 0) return 0;
    zero_rows_kernel<<<(n_rows + 7) / 8, 256, 0, er_stream(stream)>>>(n_rows, rows, row_lo, row_hi, row_offset, dst, dst_ld,
                                                                     width / 4);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_scatter_add_rows(int n_rows, const int32_t* rows, int32_t row_lo, int32_t row_hi,
                                         int32_t row_offset, const float* src, int64_t src_ld, int src_width,
                                         float* dst, int64_t dst_ld, int width, float scale, elimrec_stream_t stream) {
    ER_CHECK_ARG(width > 0 && src_width % width == 0, "src_width must be a multiple of width");
    if (n_rows <= 0) return 0;
    const int blocks = (n_rows + 7) / 8;
    scatter_add_rows_kernel<<<blocks, 256, 0, er_stream(stream)>>>(n_rows, rows, row_lo, row_hi, row_offset, src, src_ld,
                                                                   src_width / width, dst, dst_ld, width, scale);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_gather_rows(int n_rows, const int32_t* rows, const float* src, int64_t src_ld, float* dst,
                                    int64_t dst_ld, int width, elimrec_stream_t stream) {
    ER_CHECK_ARG(width % 4 == 0 && src_ld % 4 == 0 && dst_ld % 4 == 0, "width/strides must be multiples of 4");
    if (n_rows <= 0) return 0;
    gather_rows_kernel<<<(n_rows + 7) / 8, 256, 0, er_stream(stream)>>>(n_rows, rows, src, src_ld, dst, dst_ld, width / 4);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_broadcast_cols(int64_t n_rows, const float* src, int64_t src_ld, float* dst, int64_t dst_ld,
                                       int n_rep, elimrec_stream_t stream) {
    ER_CHECK_ARG(src_ld % 4 == 0 && dst_ld % 4 == 0, "strides must be multiples of 4");
    if (n_rows <= 0) return 0;
    const long long threads = n_rows * 16;
    broadcast_cols_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, er_stream(stream)>>>(n_rows, src, src_ld, dst,
                                                                                             dst_ld, n_rep);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_copy_2d(int64_t n_rows, int width, const float* src, int64_t src_ld, float* dst, int64_t dst_ld,
                                elimrec_stream_t stream) {
    ER_CHECK_ARG(width % 4 == 0 && src_ld % 4 == 0 && dst_ld % 4 == 0, "width/strides must be multiples of 4");
    if (n_rows <= 0) return 0;
    const long long threads = n_rows * (width / 4);
    copy_2d_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, er_stream(stream)>>>(n_rows, width / 4, src, src_ld, dst,
                                                                                     dst_ld);
    ER_LAUNCH_CHECK();
    return 0;
}

