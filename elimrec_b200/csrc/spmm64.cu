// 64-wide propagation, both CSR halves in ONE launch - the hot SpMM of the linear schedule (csrc/linsched.cu):
// p_k = A_hat p_{k-1} (models/EliMRec.py:244 for the id graph; by linearity also the E_u part of the modality graphs).
//
// Why not spmm_seg_kernel<64>: with one warp per row a 64-float row (256 B) keeps only half a warp busy per edge, and every
// row pays three DEPENDENT round trips to L2 (segment descriptor -> col/val -> neighbour rows) before its first FMA; at an
// average degree of 8-17 that chain, not bandwidth, set the time (6 TB/s L2->SM against ~11 TB/s for the 256-wide launch),
// and the split rows were a second launch that queued behind the first.
// Here an 8-lane GROUP owns a work item (a lane holds 2 x float4 = columns [4l, 4l+4) and [32+4l, 32+4l+4)), so a warp works
// on FOUR items at once: the dependent chain is shared by 4 rows, a warp-level LDG.128 still covers whole 128-byte lines,
// and UNR edges per group (2 x UNR LDG.128 per lane) are in flight.  Work items (elimrec_b200/graph.py build_segments64):
// rows of <= 64 edges are one item, sorted by descending degree so the four rows of a warp run in lock step; longer rows are
// dealt evenly over several items that come FIRST in the list, write partial sums to scratch, and the last-arriving item of
// the row (atomic counter) adds the partials in item order - deterministic, no float atomics, and the reduction of the
// hottest rows overlaps the bulk of the launch.  The two halves (user rows <- item rows, item rows <- user rows) are
// independent and share the grid: one launch per propagation layer.
// Edges are accumulated in CSR order: a masked launch gives the bits of the dense one.
#include "common.cuh"

namespace {

struct Half64 {
    const int4* item;      // {row, edge_begin, edge_end, split_row_id or -1}; split-row items first
    int n_item;
    const int2* hrow;      // split row h: {first item, number of items}
    int* counter;          // per split row, zero on entry, left zero
    float* partial;        // [n_split_items x W]
    const int* col;
    const float* val;
    const float* X;
    long long ldx;
    float* Y;
    long long ldy;
    const unsigned char* row_mask;   // NULL: every row.  else rows with 0 are skipped (output untouched)
    const unsigned char* col_mask;   // NULL: every edge. else edges to columns with 0 are dropped before the gather
    const float* addend;             // NULL, or Y[row] += addend[row] on the rows add_mask marks (NULL: every row)
    long long ld_add;
    const unsigned char* add_mask;
    // fused Adam (last hop of the backward chain: the finished row IS d loss / d table[row]): update the [rows x 64] table
    // in place instead of storing the gradient; old_out (may be NULL) receives the parameter row as it was before the update
    float* adam_p;
    float* adam_m;
    float* adam_v;
    float* adam_old;
};

struct AdamC {
    const double* consts;            // device: {lr / bias_correction1, sqrt(bias_correction2)} (elimrec_adam_tick)
    float b2, omb1, omb2, eps, wd;
};

// The slice of a W = 8 * FPL wide row that one lane of an 8-lane group holds.  FPL = 8: columns [4l, 4l+4) and [32+4l, 32+4l+4)
// (two 128-byte lines per row, a warp-level LDG.128 covers whole lines); FPL = 4 / 2 / 1: columns [FPL*l, FPL*(l+1)) - the
// column-sharded multi-GPU mode, where a rank propagates 64 / world columns of every row.
template <int FPL>
struct RowVec {
    float f[FPL];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < FPL; ++i) f[i] = 0.f;
    }
    template <bool CG>
    __device__ __forceinline__ void load(const float* row, int gl) {
        if (FPL == 8) {
            const float4* p = reinterpret_cast<const float4*>(row) + gl;
            const float4 a = CG ? __ldcg(p) : __ldg(p), b = CG ? __ldcg(p + 8) : __ldg(p + 8);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4 % FPL] = b.x; f[5 % FPL] = b.y; f[6 % FPL] = b.z; f[7 % FPL] = b.w;
        } else if (FPL == 4) {
            const float4* p = reinterpret_cast<const float4*>(row) + gl;
            const float4 a = CG ? __ldcg(p) : __ldg(p);
            f[0] = a.x; f[1 % FPL] = a.y; f[2 % FPL] = a.z; f[3 % FPL] = a.w;
        } else if (FPL == 2) {
            const float2* p = reinterpret_cast<const float2*>(row) + gl;
            const float2 a = CG ? __ldcg(p) : __ldg(p);
            f[0] = a.x; f[1 % FPL] = a.y;
        } else {
            f[0] = CG ? __ldcg(row + gl) : __ldg(row + gl);
        }
    }
    __device__ __forceinline__ void load_plain(const float* row, int gl) {      // read-write memory (Adam state)
        if (FPL == 8) {
            const float4* p = reinterpret_cast<const float4*>(row) + gl;
            const float4 a = p[0], b = p[8];
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4 % FPL] = b.x; f[5 % FPL] = b.y; f[6 % FPL] = b.z; f[7 % FPL] = b.w;
        } else if (FPL == 4) {
            const float4 a = reinterpret_cast<const float4*>(row)[gl];
            f[0] = a.x; f[1 % FPL] = a.y; f[2 % FPL] = a.z; f[3 % FPL] = a.w;
        } else if (FPL == 2) {
            const float2 a = reinterpret_cast<const float2*>(row)[gl];
            f[0] = a.x; f[1 % FPL] = a.y;
        } else {
            f[0] = row[gl];
        }
    }
    __device__ __forceinline__ void store(float* row, int gl) const {
        if (FPL == 8) {
            float4* p = reinterpret_cast<float4*>(row) + gl;
            p[0] = make_float4(f[0], f[1], f[2], f[3]);
            p[8] = make_float4(f[4 % FPL], f[5 % FPL], f[6 % FPL], f[7 % FPL]);
        } else if (FPL == 4) {
            reinterpret_cast<float4*>(row)[gl] = make_float4(f[0], f[1 % FPL], f[2 % FPL], f[3 % FPL]);
        } else if (FPL == 2) {
            reinterpret_cast<float2*>(row)[gl] = make_float2(f[0], f[1 % FPL]);
        } else {
            row[gl] = f[0];
        }
    }
    __device__ __forceinline__ void fma(float w, const RowVec& v) {
#pragma unroll
        for (int i = 0; i < FPL; ++i) f[i] = fmaf(w, v.f[i], f[i]);
    }
    __device__ __forceinline__ void add(const RowVec& v) {
#pragma unroll
        for (int i = 0; i < FPL; ++i) f[i] += v.f[i];
    }
};

// torch.optim.Adam single-tensor math, as in adam_multi_kernel (csrc/bpr.cu)
template <int FPL>
__device__ __forceinline__ void adam_vec(RowVec<FPL>& p, const RowVec<FPL>& g, RowVec<FPL>& m, RowVec<FPL>& v, const AdamC& c,
                                         float step_size, float bc2s) {
#pragma unroll
    for (int i = 0; i < FPL; ++i) {
        const float gr = fmaf(c.wd, p.f[i], g.f[i]);
        m.f[i] = m.f[i] + (gr - m.f[i]) * c.omb1;
        v.f[i] = fmaf(c.omb2 * gr, gr, v.f[i] * c.b2);
        p.f[i] = p.f[i] - step_size * (m.f[i] / (sqrtf(v.f[i]) / bc2s + c.eps));
    }
}

// COMPACT (launches with a row mask): the CTA first looks at its 32 items, keeps the marked ones (order preserved) and deals
// them to its first groups - warps left without work exit, so a 50 % mask costs half the time instead of all of it (the four
// items of a warp would otherwise finish only when the slowest unmasked one does).
// ADAM (fused optimizer epilogue): Adam of a row right after its gradient is complete, on the registers that hold it.
// PREF fetches the table / moment rows BEFORE the gather loop (their HBM latency then hides under the L2 gathers, at 104
// registers and 2 CTAs per SM); without it they are fetched after the loop at 64 registers and 4 CTAs per SM - measured on the
// Tiktok-shape graph (tools/spmm64_bench.py): 76.9 us with, 59.1 us without (the shipped form), 58.1 us at <UNR 2, 4 CTAs>;
// variants that spill to reach 5 CTAs per SM take 70-83 us.
template <int FPL, int UNR, int MINB, bool COMPACT, bool ADAM, int PREF = 1>
__global__ void __launch_bounds__(256, MINB)
spmm64_pair_kernel(const Half64 a, const Half64 b, const AdamC adam) {
    constexpr int W = 8 * FPL;
    typedef RowVec<FPL> Vec;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, gl = lane & 7;
    long long it_all = (((long long)blockIdx.x * 256 + threadIdx.x) >> 5) * 4 + (lane >> 3);
    bool active = true;
    if (COMPACT) {
        __shared__ long long list[32];
        __shared__ int n_list;
        if (threadIdx.x < 32) {
            const long long mine_it = (long long)blockIdx.x * 32 + threadIdx.x;
            const bool ia = mine_it < a.n_item;
            const long long ix = ia ? mine_it : mine_it - a.n_item;
            bool keep = ix < (ia ? a.n_item : b.n_item);
            if (keep) {
                const unsigned char* rm = ia ? a.row_mask : b.row_mask;
                if (rm != nullptr) keep = __ldg(rm + __ldg((ia ? a.item : b.item) + ix).x) != 0;
            }
            const unsigned bal = __ballot_sync(full, keep);
            if (keep) list[__popc(bal & ((1u << lane) - 1u))] = mine_it;
            if (lane == 0) n_list = __popc(bal);
        }
        __syncthreads();
        const int slot = threadIdx.x >> 3;                 // group index inside the CTA
        if ((slot & ~3) >= n_list) return;                 // the whole warp is beyond the compacted list
        active = slot < n_list;
        it_all = active ? list[slot] : 0;
    }
    const bool in_a = it_all < a.n_item;
#define SEL(f) (in_a ? a.f : b.f)                  /* kernel parameters: a select between two constant-bank values */
    const long long idx = in_a ? it_all : it_all - a.n_item;
    const int* col = SEL(col);
    const float* val = SEL(val);
    const float* X = SEL(X);
    const long long ldx = SEL(ldx);
    const unsigned char* cmask = SEL(col_mask);
    active = active && idx < SEL(n_item);
    int4 sg = make_int4(0, 0, 0, -1);
    if (active) {
        sg = __ldg(SEL(item) + idx);
        if (!COMPACT) {
            const unsigned char* rmask = SEL(row_mask);
            if (rmask != nullptr && __ldg(rmask + sg.x) == 0) active = false;
        }
    }
    // fused Adam: table / moment rows of a whole-row item, in flight while the neighbours are gathered
    Vec pv, mv, vv;
    float* ap = nullptr;
    if (ADAM) {
        ap = SEL(adam_p);
        if (PREF == 2 && active && ap != nullptr && sg.w < 0) {      // no registers held: the rows are only pulled into L2
            const long long o = (long long)sg.x * W + gl * (FPL >= 4 ? 4 : FPL);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ap + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(SEL(adam_m) + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(SEL(adam_v) + o));
        }
        if (PREF == 1 && active && ap != nullptr && sg.w < 0) {
            const long long o = (long long)sg.x * W;
            pv.load_plain(ap + o, gl);
            mv.load_plain(SEL(adam_m) + o, gl);
            vv.load_plain(SEL(adam_v) + o, gl);
        }
    }
    const int beg = active ? sg.y : 0, end = active ? sg.z : 0;
    int n_it = (end - beg + 7) >> 3;
    n_it = __reduce_max_sync(full, n_it);       // shuffles below need all four groups in step

    Vec acc;
    acc.zero();
    for (int it = 0; it < n_it; ++it) {
        const int e = beg + it * 8 + gl;
        int c = 0;
        float w = 0.f;
        bool live = e < end;
        if (live) {
            c = __ldg(col + e);
            w = __ldg(val + e);
            if (cmask != nullptr && __ldg(cmask + c) == 0) live = false;
        }
        const unsigned alive = __ballot_sync(full, live);      // bit per lane: edge exists and its column is marked
        const unsigned mine = (alive >> (lane & 24)) & 0xffu;  // the 8 bits of this group
#pragma unroll
        for (int u0 = 0; u0 < 8; u0 += UNR) {
            if (u0 > 0 && ((alive >> u0) & (0x01010101u * ((1u << (8 - u0)) - 1u))) == 0) break;   // nothing left in any group
            Vec v[UNR];
            float ww[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int cc = __shfl_sync(full, c, u0 + u, 8);
                const float wv = __shfl_sync(full, w, u0 + u, 8);
                const bool ok = (mine >> (u0 + u)) & 1u;
                ww[u] = ok ? wv : 0.f;
                if (ok) v[u].template load<false>(X + (long long)cc * ldx, gl);
                else v[u].zero();
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) acc.fma(ww[u], v[u]);
        }
    }

    // split rows: partial sum to scratch; the last-arriving item of the row adds all of them in item order
    const bool split = active && sg.w >= 0;
    int old = -1;
    float* partial = SEL(partial);
    int* counter = SEL(counter);
    if (split) {
        acc.store(partial + idx * W, gl);
        __threadfence();
    }
    if (__any_sync(full, split)) {
        if (split && gl == 0) old = atomicAdd(counter + sg.w, 1);
        old = __shfl_sync(full, old, 0, 8);
    }
    if (split) {
        const int2 hr = __ldg(SEL(hrow) + sg.w);
        if (old != hr.y - 1) {
            active = false;                       // somebody else finishes this row
        } else {
            __threadfence();
            acc.zero();
            const float* ps = partial + (long long)hr.x * W;
            constexpr int RU = 4;
            for (int s = 0; s < hr.y; s += RU) {
                Vec t[RU];
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    if (s + u < hr.y) t[u].template load<true>(ps + (long long)(s + u) * W, gl);
                    else t[u].zero();
                }
#pragma unroll
                for (int u = 0; u < RU; ++u) acc.add(t[u]);
            }
            if (gl == 0) counter[sg.w] = 0;     // ready for the next launch
        }
    }

    if (active) {
        const float* addend = SEL(addend);
        if (addend != nullptr) {      // the layer-mean gradient entering this layer of the backward chain (instance rows)
            const unsigned char* am = SEL(add_mask);
            if (am == nullptr || __ldg(am + sg.x) != 0) {
                Vec g;
                g.template load<false>(addend + (long long)sg.x * SEL(ld_add), gl);
                acc.add(g);
            }
        }
        if (ADAM && ap != nullptr) {
            const float step_size = (float)adam.consts[0], bc2s = (float)adam.consts[1];
            const long long o = (long long)sg.x * W;
            if (PREF != 1 || sg.w >= 0) {      // split row, finished by its last-arriving item: fetched here (PREF != 1: every row)
                pv.load_plain(ap + o, gl);
                mv.load_plain(SEL(adam_m) + o, gl);
                vv.load_plain(SEL(adam_v) + o, gl);
            }
            float* oldp = SEL(adam_old);
            if (oldp != nullptr) pv.store(oldp + o, gl);
            adam_vec<FPL>(pv, acc, mv, vv, adam, step_size, bc2s);
            pv.store(ap + o, gl);
            mv.store(SEL(adam_m) + o, gl);
            vv.store(SEL(adam_v) + o, gl);
        } else {
            acc.store(SEL(Y) + (long long)sg.x * SEL(ldy), gl);
        }
    }
#undef SEL
}

int fill_half(Half64& h, const elimrec_spmm64_half_t* s, int align, const char** err) {
    h = Half64{};
    if (s == nullptr) return 0;
    if (s->n_split_item < 0 || s->n_split_item > s->n_item) { *err = "n_split_item out of range"; return -1; }
    if (s->ldx % align != 0 || s->ldy % align != 0) { *err = "row strides must be multiples of the lane vector width"; return -1; }
    h.item = reinterpret_cast<const int4*>(s->item);
    h.n_item = s->n_item;
    h.hrow = reinterpret_cast<const int2*>(s->split_rows);
    h.counter = s->counter;
    h.partial = s->partial;
    h.col = s->col; h.val = s->val; h.X = s->X; h.ldx = s->ldx; h.Y = s->Y; h.ldy = s->ldy;
    h.row_mask = s->row_mask; h.col_mask = s->col_mask;
    h.addend = s->addend; h.ld_add = s->ld_add; h.add_mask = s->add_mask;
    if (s->addend != nullptr && s->ld_add % align != 0) { *err = "addend stride must be a multiple of the lane vector width"; return -1; }
    h.adam_p = s->adam_param; h.adam_m = s->adam_exp_avg; h.adam_v = s->adam_exp_avg_sq; h.adam_old = s->adam_old_out;
    if (h.adam_p != nullptr && (h.adam_m == nullptr || h.adam_v == nullptr)) { *err = "fused Adam needs exp_avg / exp_avg_sq"; return -1; }
    if (h.n_item > 0 && (h.X == nullptr || (h.Y == nullptr && h.adam_p == nullptr) || h.item == nullptr)) { *err = "NULL buffer"; return -1; }
    if (s->n_split_item > 0 && (h.hrow == nullptr || h.counter == nullptr || h.partial == nullptr)) {
        *err = "split rows need split_rows / counter / partial";
        return -1;
    }
    return 0;
}

}  // namespace

ELIMREC_API int elimrec_spmm64_pair(const elimrec_spmm64_half_t* a, const elimrec_spmm64_half_t* b,
                                    const elimrec_adam_consts_t* adam, int width, int variant, elimrec_stream_t stream) {
    ER_CHECK_ARG(a != nullptr, "first half required");
    ER_CHECK_ARG(width == 64 || width == 32 || width == 16 || width == 8, "width must be 64, 32, 16 or 8");
    AdamC ac{};
    const bool fused = (a->adam_param != nullptr) || (b != nullptr && b->adam_param != nullptr);
    ER_CHECK_ARG(!fused || (adam != nullptr && adam->consts_dev != nullptr), "fused Adam needs the optimizer constants");
    if (adam != nullptr) {
        ac.consts = adam->consts_dev;
        ac.b2 = (float)adam->beta2;
        ac.omb1 = (float)(1.0 - adam->beta1);      // python doubles rounded once, as torch does
        ac.omb2 = (float)(1.0 - adam->beta2);
        ac.eps = adam->eps;
        ac.wd = adam->weight_decay;
    }
    Half64 ha, hb;
    const char* err = nullptr;
    const int align = width >= 32 ? 4 : (width == 16 ? 2 : 1);
    if (fill_half(ha, a, align, &err) != 0 || fill_half(hb, b, align, &err) != 0) {
        elimrec_set_error("elimrec_spmm64_pair: %s", err);
        return -1;
    }
    const long long items = (long long)ha.n_item + hb.n_item;
    if (items <= 0) return 0;
    const unsigned blocks = (unsigned)((items + 31) / 32);      // 8 warps x 4 items per CTA
    cudaStream_t st = er_stream(stream);
    const bool compact = ha.row_mask != nullptr || hb.row_mask != nullptr;
    ER_CHECK_ARG(!(fused && compact), "fused Adam runs on the dense last hop (no row mask)");
#define PAIR_LAUNCH(FPL)                                                                                                   \
    do {                                                                                                                   \
        if (fused && variant == 1) spmm64_pair_kernel<FPL, 4, 2, false, true, 1><<<blocks, 256, 0, st>>>(ha, hb, ac);   \
        else if (fused && variant == 2) spmm64_pair_kernel<FPL, 4, 4, false, true, 2><<<blocks, 256, 0, st>>>(ha, hb, ac); \
        else if (fused) spmm64_pair_kernel<FPL, 4, 4, false, true, 0><<<blocks, 256, 0, st>>>(ha, hb, ac);             \
        else if (compact) spmm64_pair_kernel<FPL, 4, 4, true, false><<<blocks, 256, 0, st>>>(ha, hb, ac);                  \
        else if (variant == 1) spmm64_pair_kernel<FPL, 8, 2, false, false><<<blocks, 256, 0, st>>>(ha, hb, ac);            \
        else if (variant == 2) spmm64_pair_kernel<FPL, 4, 3, false, false><<<blocks, 256, 0, st>>>(ha, hb, ac);            \
        else if (variant == 4) spmm64_pair_kernel<FPL, 8, 3, false, false><<<blocks, 256, 0, st>>>(ha, hb, ac);            \
        else spmm64_pair_kernel<FPL, 4, 4, false, false><<<blocks, 256, 0, st>>>(ha, hb, ac);                              \
    } while (0)
    if (width == 64) PAIR_LAUNCH(8);
    else if (width == 32) PAIR_LAUNCH(4);
    else if (width == 16) PAIR_LAUNCH(2);
    else PAIR_LAUNCH(1);
#undef PAIR_LAUNCH
    ER_LAUNCH_CHECK();
    return 0;
}
