// Full-ranking evaluator on the tensor cores (tcgen05, kind::f16) with fp32-class accuracy.
//
// The score of (user u, item i) needs (1+M) length-64 dot products (fused table + single-modal heads,
// reference models/EliMRec.py:96-113,155-188).  To stay exact enough for top-K (neighbours in the ranking differ
// by ~1e-6) every operand is a PAIR of fp16 numbers  x*s = hi + lo  (22 significant bits, s a power of two per
// table) - the same 4 bytes per element as raw fp32 - and each dot product is three MMAs
// hi*hi + hi*lo + lo*hi  accumulated in fp32 in TMEM.
//
// Orientation: ITEMS on the TMEM lanes (MMA M = 128, streamed), USERS on the columns (MMA N = 64, resident in
// shared memory for the whole CTA).  A drain warp then sees 32 different items of ONE user per TMEM column, so
// the per-user selection is a warp ballot + a shuffle insert into a lane-distributed top-K list in registers.
//
//   warp 0      : TMA producer - ring of 4 slots, one table (hi + lo, [128 items x 64 dims] fp16, 128-byte
//                 swizzle) per slot
//   warp 1      : TMEM allocator + MMA issuer (per table: 3 terms x 4 k-steps, M=128 N=64 K=16)
//   warps 2..17 : gather the 64 user rows once (swizzled K-major layout) and the first 32 training items of every
//                 user; then drain - warp = (TMEM lane quadrant q, group of 16 user columns).
// Accumulators are double-buffered in TMEM (2 x 256 columns): the MUFU-bound epilogue of tile j overlaps the
// TMA + MMA of tile j+1.
//
// Ranking key.  Every mode's final step is a sigmoid (EliMRec.py:100-113), which is monotone, so the kernel ranks
// on its ARGUMENT and the merge kernel applies the final sigmoid to the K winners only:
//   normal : sigmoid(sigmoid(x0))                          key = x0
//   TE     : sigmoid(ui * z_1 .. z_M)                      key = 1 / prod_t (1 + e_t)            e_t = exp(-x_t)
//   TIE    : sigmoid(ui * z_1..z_M - mean_u * z_1..z_M)    key = (1 - mean_u (1 + e_0)) / prod_t (1 + e_t)
// i.e. (1+M) MUFU.EX2 + ONE MUFU.RCP per (user, item) instead of 2 (2+M) MUFU.  MODE mean computes
// mean_i sigmoid(x0) (TIE needs it first, EliMRec.py:107).
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RWAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RWAIT_DONE;\n"
        "bra RWAIT_LOOP;\n"
        "RWAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// the same box delivered to the same shared-memory offset (and signalled on the same barrier offset) of every CTA in `mask`
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// true in exactly one lane of a converged warp.  The tensor-core / TMA instructions are warp-level ("uniform") instructions:
// issued under `if (lane == 0)` ptxas wraps every one of them in its own elect-and-loop sequence (ELECT / PLOP3 / BRA.U.ANY,
// ~100 cycles of issue latency per MMA - more than the MMA itself); issued under elect.sync by a warp that runs the loop in
// step it emits the bare instruction.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrives on the barrier at this offset in every CTA of `mask` once the issuing thread's earlier MMAs have completed
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_32x4(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {   // K-major, 128-byte swizzle, 8-row groups 1024 B apart
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16: c=F32 [4,6)=1, a=F16 [7,10)=0, b=F16 [10,13)=0, both K-major, N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t IDESC_F16_128x64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

constexpr int NTMAX = 1 + ELIMREC_MAX_MODS;
constexpr int T_ITEMS = 128;                   // items per tile  = MMA M = TMEM lanes
constexpr int T_USERS = 64;                    // users per CTA   = MMA N = TMEM columns per table
constexpr int I_PART = T_ITEMS * 128;          // one [128 items x 64 dims] fp16 tile = 16 KB
constexpr int U_PART = T_USERS * 128;          // one [ 64 users x 64 dims] fp16 tile =  8 KB
constexpr int TC_CLUSTER = 2;                  // CTAs (user tiles) per cluster: every item tile is fetched from L2 once per cluster
constexpr int TC_PIECE = 2 * T_ITEMS / TC_CLUSTER;   // rows of the [hi ; lo] slot (2 x 128) that one CTA of the cluster fetches
constexpr int N_DRAIN = 16;                    // drain warps: 4 TMEM lane quadrants x 4 groups of 16 user columns
constexpr int TC_THREADS = 32 * (2 + N_DRAIN);
enum { TC_MEAN = 0, TC_NORMAL = 1, TC_TE = 2, TC_TIE = 3 };

struct RankTcMaps {
    CUtensorMap hi[NTMAX];
    CUtensorMap lo[NTMAX];
};
struct RankTcArgs {
    int n_eval, I, nt, mode, K, n_mod;
    const int* eval_users;
    const float* mean_in;
    const __half* uh[NTMAX];
    const __half* ul[NTMAX];
    float inv_scale[NTMAX];       // 1 / (user scale * item scale) of each table
    const long long* train_ptr;
    const int* train_items;
    int* part_idx;                // [n_eval][4 quadrants][K]   per-quadrant candidate lists
    float* part_val;
    double* part_sum;             // [n_eval][4]                per-quadrant row sums (MODE_MEAN)
};

// is `item` among the entries 32.. of the user's sorted train row?  (rows longer than the shared-memory cache)
__device__ __noinline__ bool train_row_tail_has(const int* eval_users, const long long* train_ptr, const int* train_items,
                                                int gu, int item) {
    const int u = __ldg(eval_users + gu);
    long long lo = __ldg(train_ptr + u) + 32;
    const long long end = __ldg(train_ptr + u + 1);
    long long hi = end;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(train_items + mid) < item) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(train_items + lo) == item;
}

// Clusters of TC_CLUSTER CTAs (user tiles) share the item stream: of every ring slot ([hi ; lo] boxes = 2 x 128 rows) each CTA
// fetches TC_PIECE rows and multicasts them into the shared memory of ALL CTAs of the cluster, so an item tile leaves L2 once
// per cluster instead of once per CTA (every CTA used to stream all (1+M) item tables: 44.6 GB per evaluation at Tiktok shape).
// A slot is refilled only when the MMA warps of every CTA have released it (the commit is multicast to all empty barriers).
template <int MODE, int NT>
__global__ void __cluster_dims__(TC_CLUSTER, 1, 1) __launch_bounds__(TC_THREADS, 1)
rank_tc_kernel(const __grid_constant__ RankTcMaps mp, const RankTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[NTMAX], empty_bar[NTMAX], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset form keeps the shared address space
    constexpr int nt = NT;
    uint8_t* smI = smem;                               // ring of NTMAX slots x [hi,lo][128 items x 128 B], one table per slot
    uint8_t* smU = smem + NTMAX * 2 * I_PART;          // [nt][hi,lo][ 64 users x 128 B]   resident
    int* smT = reinterpret_cast<int*>(smU + NTMAX * 2 * U_PART);   // [64 users][32] first training items (-1 padded)
    float* smMT = reinterpret_cast<float*>(smT + T_USERS * 32);    // [16 drain warps][mean_u x 16 | K-th key x 16]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u0 = blockIdx.x * T_USERS;
    const int n_tiles = (a.I + T_ITEMS - 1) / T_ITEMS;
    const unsigned full = 0xffffffffu;

    if (warp == 0 && lane == 0) {
        for (int b = 0; b < NTMAX; ++b) {
            mbar_init(&full_bar[b], 1);
            mbar_init(&empty_bar[b], TC_CLUSTER);      // one release per CTA of the cluster
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], N_DRAIN);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
    if (warp >= 2) {   // gather the CTA's user rows once, swizzled exactly like a SWIZZLE_128B TMA box
        const int t0 = threadIdx.x - 64;
        const int total = nt * 2 * T_USERS * 8;
        for (int idx = t0; idx < total; idx += 32 * N_DRAIN) {
            const int c = idx & 7, r = (idx >> 3) & (T_USERS - 1), part = (idx >> 9) & 1, t = idx >> 10;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (u0 + r < a.n_eval) {
                const int u = __ldg(a.eval_users + u0 + r);
                const __half* src = (part == 0 ? a.uh[t] : a.ul[t]) + (long long)u * 64;
                v = __ldg(reinterpret_cast<const uint4*>(src) + c);
            }
            *reinterpret_cast<uint4*>(smU + (t * 2 + part) * U_PART + r * 128 + ((c ^ (r & 7)) << 4)) = v;
        }
        if (MODE != TC_MEAN) {   // masking cache: membership of a candidate in the user's train row is then ONE ballot
            for (int r = warp - 2; r < T_USERS; r += N_DRAIN) {
                int v = -1;
                if (u0 + r < a.n_eval) {
                    const int u = __ldg(a.eval_users + u0 + r);
                    const long long b0 = __ldg(a.train_ptr + u), e0 = __ldg(a.train_ptr + u + 1);
                    if (b0 + lane < e0) v = __ldg(a.train_items + b0 + lane);
                }
                smT[r * 32 + lane] = v;
            }
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                          // the peer's barriers exist before anything of mine can reach them
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;
    const uint32_t crank = cluster_ctarank();
    constexpr uint16_t CMASK = (1u << TC_CLUSTER) - 1;

    if (warp == 0) {
        int it = 0;                          // running (tile, table) counter: slot = it % NTMAX
        constexpr int PPP = T_ITEMS / TC_PIECE;                // pieces per part
        const int part = (int)crank / PPP, r0 = ((int)crank % PPP) * TC_PIECE;
        for (int j = 0; j < n_tiles; ++j) {
            for (int t = 0; t < nt; ++t, ++it) {
                const int s = it & (NTMAX - 1);
                mbar_wait(&empty_bar[s], ((it / NTMAX) & 1) ^ 1);      // released by EVERY CTA of the cluster
                if (elect_one()) {
                    mbar_expect_tx(&full_bar[s], (uint32_t)(2 * I_PART));  // my piece + the peers'
                    tma_load_2d_mc(part == 0 ? &mp.hi[t] : &mp.lo[t], &full_bar[s], smI + (s * 2 + part) * I_PART + r0 * 128, 0,
                                   j * T_ITEMS + r0, CMASK);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // the whole warp walks the loop (waits included) so that every address below is warp-uniform; one elected lane issues
        int it = 0;
        for (int j = 0; j < n_tiles; ++j) {
            const int buf = j & 1;
            mbar_wait(&tempty_bar[buf], ((j >> 1) & 1) ^ 1);
            for (int t = 0; t < nt; ++t, ++it) {
                const int s = it & (NTMAX - 1);
                mbar_wait(&full_bar[s], (it / NTMAX) & 1);
                tc_fence_after();
                const uint32_t ih = smem_u32(smI + (s * 2) * I_PART), il = ih + I_PART;
                const uint32_t uh = smem_u32(smU + (t * 2) * U_PART), ul = uh + U_PART;
                const uint32_t d = tmem_d + buf * 256 + t * 64;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {   // K = 64 = 4 x 16 halves (32 bytes inside the 128-byte atom)
                        umma_f16(d, umma_desc_sw128(il + k * 32), umma_desc_sw128(uh + k * 32), IDESC_F16_128x64, k != 0);
                        umma_f16(d, umma_desc_sw128(ih + k * 32), umma_desc_sw128(ul + k * 32), IDESC_F16_128x64, 1);
                        umma_f16(d, umma_desc_sw128(ih + k * 32), umma_desc_sw128(uh + k * 32), IDESC_F16_128x64, 1);
                    }
                    umma_commit_mc(&empty_bar[s], CMASK);  // ring slot consumed here: tell both producers
                    if (t == nt - 1) umma_commit(&tfull_bar[buf]);   // accumulators of tile j complete
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;                 // TMEM lane quadrant: items 32q .. 32q+31 of every tile
        const int cg = (warp - 2) >> 2;         // column group: users 16cg .. 16cg+15 of this CTA
        const int K = a.K;
        float nls[NT];                          // -log2(e) / (user scale * item scale): e_t = 2^(acc_t * nls_t)
#pragma unroll
        for (int t = 0; t < NT; ++t) nls[t] = -1.4426950408889634f * a.inv_scale[t];
        if (MODE == TC_MEAN) {
            float psum[16];                     // per-lane fp32 partial sums, flushed into fp64 every 8 tiles
            double usum[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) { psum[c] = 0.f; usum[c] = 0.0; }
            for (int j = 0; j < n_tiles; ++j) {
                const int buf = j & 1;
                mbar_wait(&tfull_bar[buf], (j >> 1) & 1);
                tc_fence_after();
                float acc[16];
                tmem_ld_32x16(tmem_d + ((uint32_t)(q * 32) << 16) + buf * 256 + cg * 16, acc);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[buf]);
                const bool valid = j * T_ITEMS + q * 32 + lane < a.I;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    float e, r;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(acc[c] * nls[0]));
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
                    psum[c] += valid ? r : 0.f;
                }
                if ((j & 7) == 7 || j == n_tiles - 1) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        float p = psum[c];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(full, p, o);
                        usum[c] += (double)p;
                        psum[c] = 0.f;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const int gu = u0 + cg * 16 + c;
                if (gu < a.n_eval && lane == 0) a.part_sum[(long long)gu * 4 + q] = usum[c];
            }
        } else {
            float lv[16];                       // lane e holds entry e of the key-sorted top-K list of user column c
            int li[16];                         // (this quadrant's items only; the merge kernel joins the 4 quadrants)
            float* myMean = smMT + (warp - 2) * 32;
            float* myThr = myMean + 16;         // current K-th key per column; +inf for the padding columns of the last CTA
            unsigned long_mask = 0;             // columns whose train row continues past the 32-entry cache
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                lv[c] = -INFINITY; li[c] = -1;
                if (smT[(cg * 16 + c) * 32 + 31] >= 0) long_mask |= 1u << c;
            }
            if (lane < 16) {
                const int gu = u0 + cg * 16 + lane;
                myMean[lane] = (MODE == TC_TIE && gu < a.n_eval) ? __ldg(a.mean_in + gu) : 0.f;
                myThr[lane] = (gu < a.n_eval) ? -INFINITY : INFINITY;
            }
            __syncwarp();
            for (int j = 0; j < n_tiles; ++j) {
                const int buf = j & 1;
                mbar_wait(&tfull_bar[buf], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t tb = tmem_d + ((uint32_t)(q * 32) << 16) + buf * 256 + cg * 16;
                const int item0 = j * T_ITEMS + q * 32;
                const bool valid = item0 + lane < a.I;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {   // 4 user columns at a time: ILP 4 through the MUFU chain, one branch
                    float acc[NT][4];
#pragma unroll
                    for (int t = 0; t < NT; ++t) tmem_ld_32x4(tb + t * 64 + qd * 4, acc[t]);
                    tmem_ld_wait();
                    if (qd == 3) {              // last TMEM read of this buffer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[buf]);
                    }
                    const float4 thr4 = *reinterpret_cast<const float4*>(myThr + qd * 4);
                    const float thr_[4] = {thr4.x, thr4.y, thr4.z, thr4.w};
                    float mean_[4] = {0.f, 0.f, 0.f, 0.f};
                    if (MODE == TC_TIE) {
                        const float4 m4 = *reinterpret_cast<const float4*>(myMean + qd * 4);
                        mean_[0] = m4.x; mean_[1] = m4.y; mean_[2] = m4.z; mean_[3] = m4.w;
                    }
                    float key[4];
                    unsigned m[4];
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        if (MODE == TC_NORMAL) {
                            key[c4] = acc[0][c4];
                        } else {
                            float e0;
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(acc[0][c4] * nls[0]));
                            const float p0 = 1.f + e0;
                            float den = p0;
#pragma unroll
                            for (int t = 1; t < NT; ++t) {
                                float e;
                                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(acc[t][c4] * nls[t]));
                                den *= 1.f + e;
                            }
                            float r;
                            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
                            key[c4] = (MODE == TC_TE) ? r : fmaf(-mean_[c4], p0, 1.f) * r;
                        }
                        m[c4] = __ballot_sync(full, valid && key[c4] > thr_[c4]);
                    }
                    if ((m[0] | m[1] | m[2] | m[3]) == 0u) continue;
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        unsigned mm = m[c4];
                        if (mm == 0u) continue;
                        const int c = qd * 4 + c4;
                        const int tcache = smT[(cg * 16 + c) * 32 + lane];   // first 32 training items of this user
                        float thr = thr_[c4];
                        while (mm) {
                            const int b = __ffs(mm) - 1;
                            mm &= mm - 1;
                            const float cv = __shfl_sync(full, key[c4], b);
                            if (!(cv > thr)) continue;
                            const int cid = item0 + b;
                            // training item? (uni_evaluator.py:149-154) - one ballot against the cached row; rows longer
                            // than 32 items continue with a binary search in global memory (rare, out of line)
                            if (__ballot_sync(full, tcache == cid) != 0u) continue;
                            if (((long_mask >> c) & 1u) &&
                                train_row_tail_has(a.eval_users, a.train_ptr, a.train_items, u0 + cg * 16 + c, cid))
                                continue;
                            const int pos = __popc(__ballot_sync(full, lane < K && lv[c] >= cv));
                            const float pv = __shfl_up_sync(full, lv[c], 1);
                            const int pi = __shfl_up_sync(full, li[c], 1);
                            if (lane == pos) { lv[c] = cv; li[c] = cid; }
                            else if (lane > pos) { lv[c] = pv; li[c] = pi; }
                            thr = __shfl_sync(full, lv[c], K - 1);
                        }
                        if (lane == 0) myThr[c] = thr;
                    }
                    __syncwarp();
                }
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const int gu = u0 + cg * 16 + c;
                if (gu < a.n_eval && lane < K) {
                    a.part_idx[((long long)gu * 4 + q) * K + lane] = li[c];
                    a.part_val[((long long)gu * 4 + q) * K + lane] = lv[c];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, 512);
    cluster_sync_all();                          // no CTA leaves while the peer's releases / boxes can still land in it
}

constexpr int RANK_TC_SMEM = NTMAX * 2 * I_PART + NTMAX * 2 * U_PART + T_USERS * 32 * 4 + N_DRAIN * 16 * 8 + 1024;

// one warp per user: merge the 4 per-quadrant lists (score descending, index ascending) / add the 4 row sums
__global__ void rank_tc_merge_kernel(int n_eval, int K, int I, int what, float inv_scale0, const int* __restrict__ part_idx,
                                     const float* __restrict__ part_val, const double* __restrict__ part_sum,
                                     int* __restrict__ out_idx, float* __restrict__ out_val, float* __restrict__ out_mean) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    if (r >= n_eval) return;
    if (what == TC_MEAN) {
        if (lane == 0) {
            const double* p = part_sum + (long long)r * 4;
            out_mean[r] = (float)((((p[0] + p[1]) + p[2]) + p[3]) / (double)I);
        }
        return;
    }
    float lv = -INFINITY;
    int li = INT_MAX;
    for (int e = 0; e < 4 * K; ++e) {
        const float cv = __ldg(part_val + (long long)r * 4 * K + e);
        const int cid = __ldg(part_idx + (long long)r * 4 * K + e);
        if (cid < 0) continue;                   // unused slot (warp-uniform)
        const int pos = __popc(__ballot_sync(full, lane < K && (lv > cv || (lv == cv && li < cid))));
        if (pos >= K) continue;
        const float pv = __shfl_up_sync(full, lv, 1);
        const int pi = __shfl_up_sync(full, li, 1);
        if (lane == pos) { lv = cv; li = cid; }
        else if (lane > pos) { lv = pv; li = pi; }
    }
    if (lane < K) {      // key -> score: the final sigmoid(s) of the mode, applied to the K winners only
        float v = lv;
        if (li != INT_MAX) {
            if (what == TC_NORMAL) v = 1.f / (1.f + expf(-(v * inv_scale0)));
            v = 1.f / (1.f + expf(-v));
        }
        out_idx[(long long)r * K + lane] = (li == INT_MAX) ? -1 : li;
        out_val[(long long)r * K + lane] = v;
    }
}

__global__ void split_fp16_kernel(long long n, const float* __restrict__ src, float scale, __half* __restrict__ hi,
                                  __half* __restrict__ lo) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = src[i] * scale;        // scale is a power of two: exact
    const __half h = __float2half_rn(x);
    hi[i] = h;
    lo[i] = __float2half_rn(x - __half2float(h));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
int make_map_f16(CUtensorMap* map, const void* base, int64_t rows) {   // [rows x 64] fp16, box [64 dims x TC_PIECE items]
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) return -1;
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64u, (cuuint32_t)TC_PIECE};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

template <int MODE, int NT>
int launch_rank_tc(const RankTcMaps& mp, const RankTcArgs& a, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(rank_tc_kernel<MODE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, RANK_TC_SMEM);
        if (e != cudaSuccess) {
            elimrec_set_error("rank_tc: shared-memory opt-in (%d B) failed: %s", RANK_TC_SMEM, cudaGetErrorString(e));
            return -3;
        }
        configured = true;
    }
    const int blocks = ((a.n_eval + T_USERS - 1) / T_USERS + TC_CLUSTER - 1) / TC_CLUSTER * TC_CLUSTER;   // padding CTAs rank nothing
    rank_tc_kernel<MODE, NT><<<blocks, TC_THREADS, RANK_TC_SMEM, st>>>(mp, a);
    return 0;
}

}  // namespace

ELIMREC_API int elimrec_split_fp16(int64_t n, const float* src, float scale, void* hi, void* lo, elimrec_stream_t stream) {
    if (n <= 0) return 0;
    split_fp16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, er_stream(stream)>>>(n, src, scale, (__half*)hi, (__half*)lo);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_rank_tc(const elimrec_rank_tc_tables_t* t, int what, int n_eval, const int32_t* eval_users,
                                const float* ui_mean, const int64_t* train_ptr, const int32_t* train_items, int K,
                                int32_t* topk_idx, float* topk_val, float* mean_out, void* workspace,
                                elimrec_stream_t stream) {
    ER_CHECK_ARG(t != nullptr && t->n_mod >= 0 && t->n_mod <= ELIMREC_MAX_MODS && t->mode >= 0 && t->mode <= 2, "bad descriptor");
    ER_CHECK_ARG(what == 0 || (K >= 1 && K <= 32), "K must be in [1, 32]");
    ER_CHECK_ARG(what == 0 || t->mode != 2 || ui_mean != nullptr, "TIE needs ui_mean");
    if (n_eval <= 0) return 0;
    RankTcMaps mp;
    RankTcArgs a{};
    a.n_eval = n_eval; a.I = t->num_items; a.mode = t->mode; a.K = K; a.n_mod = (t->mode == 0) ? 0 : t->n_mod;
    a.nt = 1 + a.n_mod;
    a.eval_users = eval_users; a.mean_in = ui_mean;
    a.train_ptr = (const long long*)train_ptr; a.train_items = train_items;
    ER_CHECK_ARG(workspace != nullptr, "workspace required (elimrec_rank_tc_workspace_bytes)");
    // workspace: [n_eval x 4] doubles | [n_eval x 4 x K] floats | [n_eval x 4 x K] ints
    a.part_sum = reinterpret_cast<double*>(workspace);
    a.part_val = reinterpret_cast<float*>(a.part_sum + (size_t)n_eval * 4);
    a.part_idx = reinterpret_cast<int*>(a.part_val + (size_t)n_eval * 4 * 32);
    int bad = 0;
    for (int i = 0; i < NTMAX; ++i) {
        const int s = i < a.nt ? i : 0;
        a.uh[i] = (const __half*)t->user_hi[s]; a.ul[i] = (const __half*)t->user_lo[s];
        a.inv_scale[i] = t->inv_scale[s];
        bad |= make_map_f16(&mp.hi[i], t->item_hi[s], t->num_items) | make_map_f16(&mp.lo[i], t->item_lo[s], t->num_items);
    }
    if (bad) {
        elimrec_set_error("elimrec_rank_tc: cuTensorMapEncodeTiled failed");
        return -3;
    }
    cudaStream_t st = er_stream(stream);
    int rc, kind = TC_MEAN;
    if (what == 0) {
        rc = launch_rank_tc<TC_MEAN, 1>(mp, a, st);
    } else if (t->mode == 0) {
        kind = TC_NORMAL;
        rc = launch_rank_tc<TC_NORMAL, 1>(mp, a, st);
    } else if (t->mode == 1) {
        kind = TC_TE;
        rc = a.nt == 1 ? launch_rank_tc<TC_TE, 1>(mp, a, st) : a.nt == 2 ? launch_rank_tc<TC_TE, 2>(mp, a, st)
           : a.nt == 3 ? launch_rank_tc<TC_TE, 3>(mp, a, st) : launch_rank_tc<TC_TE, 4>(mp, a, st);
    } else {
        kind = TC_TIE;
        rc = a.nt == 1 ? launch_rank_tc<TC_TIE, 1>(mp, a, st) : a.nt == 2 ? launch_rank_tc<TC_TIE, 2>(mp, a, st)
           : a.nt == 3 ? launch_rank_tc<TC_TIE, 3>(mp, a, st) : launch_rank_tc<TC_TIE, 4>(mp, a, st);
    }
    if (rc != 0) return rc;
    ER_LAUNCH_CHECK();
    rank_tc_merge_kernel<<<(n_eval + 7) / 8, 256, 0, st>>>(n_eval, K, t->num_items, kind, a.inv_scale[0], a.part_idx, a.part_val,
                                                                         a.part_sum, topk_idx, topk_val, mean_out);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int64_t elimrec_rank_tc_workspace_bytes(int n_eval) {
    return (int64_t)n_eval * 4 * 8 + (int64_t)n_eval * 4 * 32 * 4 * 2;
}
