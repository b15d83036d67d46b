// Tensor-core linear layers for sm_100a: tcgen05.mma (kind::tf32) with TMA-fed shared-memory operands
// and the fp32 accumulator in TMEM.  Replaces cuBLAS SGEMM behind nn.Linear for the modal-feature
// projections v_dense / a_dense / t_dense (reference models/EliMRec.py:233-236), which stream the
// constant [I x D_m] feature matrices every step and are HBM-bound (32 flop per feature byte).
//
//   fwd  : Y[M x 64] (ldy) = X[M x K] (ldx) * W[64 x K]^T + b
//          CTA = one 128-row tile of X; per k-block of 32 floats (one 128-byte swizzle atom) TMA brings
//          a [128 x 32] tile of X and a [64 x 32] tile of W (both K-major, SWIZZLE_128B), one elected
//          thread issues 4 x tcgen05.mma 128x64x8; a 4-stage mbarrier ring overlaps TMA with MMA; four
//          epilogue warps read the accumulator with tcgen05.ld, add the bias and write coalesced rows.
//   wgrad: dW[64 x K] = dY[M x 64]^T * X[M x K]   (reduction over the M rows)
//          computed transposed, D^T[c, n] = sum_m X[m, c] dY[m, n], so that both operands are
//          MN-major tiles straight from row-major memory: A' = X^T (128 columns of X per MMA tile),
//          B' = dY.  A CTA owns up to 256 columns of X and a contiguous range of rows, keeps its
//          [256 x 64] partial in TMEM for the whole range and writes it once; a second kernel
//          reduces the row-range partials in a fixed order (deterministic).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2-5 = epilogue (warp_id % 4 selects the TMEM lane quadrant a warp may read).
#include <cuda.h>
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// true in exactly one lane of a converged warp.  TMA / tensor-core instructions are warp-level ("uniform") instructions: under
// `if (lane == 0)` ptxas wraps each of them in an elect-and-loop sequence (ELECT / PLOP3 / BRA.U.ANY, ~100 cycles of issue
// latency per MMA - more than a 128x64 MMA itself takes); under elect.sync, with the whole warp walking the loop in step so
// that every operand is warp-uniform, it emits the bare predicated instruction.  The producer and MMA warps below therefore
// run their loops (barrier waits included) on all 32 lanes and the issuing primitives elect their lane themselves.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {      // one elected lane
    if (elect_one())
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {   // one elected lane
    if (elect_one())
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate; issued by one elected lane of a converged warp
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (elect_one())
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {      // one elected lane (the one that issued the MMAs)
    if (elect_one())
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (32*(warp%4) + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor, version 1):
//   [0,14) start>>4 | [16,30) leading byte offset>>4 | [32,46) stride byte offset>>4 | [46,48) version=1 | [61,64) layout=2
//   layout: 2 = SWIZZLE_128B (16-byte swizzle atoms), 1 = SWIZZLE_128B_BASE32B (32-byte atoms; the only layout the
//   hardware accepts for MN-major 32-bit (tf32) operands - cutlass sm100_common.inl "for mn-major tf32 operands,
//   SW128_32B is the only available smem layout"; TMA counterpart CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return umma_desc(saddr, lbo_bytes, sbo_bytes, 2);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=TF32 [7,10), b=TF32 [10,13),
// a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent: one CTA per SM walks row tiles  tile = blockIdx.x, + gridDim.x, ...  with ONE TMA ring running across tiles
// (6 stages x 24 KB in flight per SM covers HBM latency at the SM's share of the bandwidth) and two TMEM accumulators,
// so the store epilogue of tile j overlaps the loads + MMAs of tile j+1.  The grid never exceeds the SM count: an
// L2-bound kernel launched next to it (the narrow layer-1 SpMM, models/EliMRec.py:244) finds free warp slots on every SM
// instead of queueing behind a many-wave grid, and there is no wave-quantisation tail.
constexpr int F_BM = 128, F_BN = 64, F_BK = 32, F_STAGES = 6;
constexpr int F_A_BYTES = F_BM * F_BK * 4, F_B_BYTES = F_BN * F_BK * 4, F_STAGE = F_A_BYTES + F_B_BYTES;
constexpr int F_STG = 4 * 32 * 65 * 4;                       // one 32x64 (+pad) transpose tile per epilogue warp
constexpr int F_SMEM = F_STAGES * F_STAGE + F_STG + 1024;    // + slack for the 1024-byte alignment swizzle-128B needs
constexpr int F_TMEM_COLS = 128;                             // 2 accumulators x 64 columns

constexpr int F_MAX_PROB = 4;       // problems (modalities) per launch
struct FwdMulti {
    CUtensorMap A[F_MAX_PROB], B[F_MAX_PROB];
    const float* bias[F_MAX_PROB];
    float* Y[F_MAX_PROB];
    long long ldy[F_MAX_PROB];
    int M[F_MAX_PROB], num_kb[F_MAX_PROB];
    int tile_start[F_MAX_PROB + 1];  // problem p owns tiles [tile_start[p], tile_start[p+1])
    int n;
};
__device__ __forceinline__ int fwd_problem_of(const FwdMulti& mp, int tile) {
    int p = 0;
    while (p + 1 < mp.n && tile >= mp.tile_start[p + 1]) ++p;
    return p;
}

// Several projections (one per modality: same 64 output columns each, different K) share ONE launch: the tile list is the
// concatenation of the problems' row tiles, largest K first, dealt round-robin to the CTAs.
__global__ void __launch_bounds__(192, 1)
linear_tf32_fwd_kernel(const __grid_constant__ FwdMulti mp) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[F_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[F_STAGES];
    __shared__ __align__(8) uint64_t tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int n_tiles = mp.tile_start[mp.n];
    if (warp == 0 && lane == 0) {
        for (int p = 0; p < mp.n; ++p) {
            tma_prefetch_desc(&mp.A[p]);
            tma_prefetch_desc(&mp.B[p]);
        }
        for (int s = 0; s < F_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, F_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;

    if (warp == 0) {
        {   // whole warp in step, one elected lane issues (elect_one)
            int it = 0;                          // k-blocks issued so far (ring position), across tiles
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int p = fwd_problem_of(mp, tile);
                const int m0 = (tile - mp.tile_start[p]) * F_BM, num_kb = mp.num_kb[p];
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % F_STAGES;
                    mbar_wait(&empty_bar[s], ((it / F_STAGES) & 1) ^ 1);
                    mbar_expect_tx(&full_bar[s], F_STAGE);
                    uint8_t* a = smem + s * F_STAGE;
                    tma_load_2d(&mp.A[p], &full_bar[s], a, kb * F_BK, m0);
                    tma_load_2d(&mp.B[p], &full_bar[s], a + F_A_BYTES, kb * F_BK, 0);
                }
            }
        }
    } else if (warp == 1) {
        {   // whole warp in step, one elected lane issues (elect_one)
            constexpr uint32_t idesc = umma_idesc_tf32(F_BM, F_BN, 0, 0);
            int it = 0, j = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
                const int buf = j & 1;
                mbar_wait(&tempty_bar[buf], ((j >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem_d + buf * F_BN;
                const int num_kb = mp.num_kb[fwd_problem_of(mp, tile)];
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % F_STAGES;
                    mbar_wait(&full_bar[s], (it / F_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a = smem_u32(smem + s * F_STAGE);
                    const uint32_t b = a + F_A_BYTES;
#pragma unroll
                    for (int k = 0; k < F_BK / 8; ++k) {  // UMMA_K = 8 tf32 = 32 bytes inside the 128-byte atom
                        const uint64_t da = umma_desc_sw128(a + k * 32, 16, 1024);
                        const uint64_t db = umma_desc_sw128(b + k * 32, 16, 1024);
                        umma_tf32(d, da, db, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[s]);  // smem slot reusable once these MMAs have read it
                }
                umma_commit(&tfull_bar[buf]);    // accumulator of this tile complete
            }
        }
    } else {
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        float* st = reinterpret_cast<float*>(smem + F_STAGES * F_STAGE) + q * (32 * 65);
        int j = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
            const int buf = j & 1;
            const int p = fwd_problem_of(mp, tile);
            const float* bias = mp.bias[p];
            float* Y = mp.Y[p];
            const long long ldy = mp.ldy[p];
            const int M = mp.M[p];
            float bj[2] = {0.f, 0.f};
            if (bias != nullptr) { bj[0] = __ldg(bias + lane); bj[1] = __ldg(bias + 32 + lane); }
            mbar_wait(&tfull_bar[buf], (j >> 1) & 1);
            tc_fence_after();
            float v[64];
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + buf * F_BN;
            tmem_ld_32x32(taddr, v);
            tmem_ld_32x32(taddr + 32, v + 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
            // transpose through shared memory for coalesced row stores
#pragma unroll
            for (int c = 0; c < 64; ++c) st[lane * 65 + c] = v[c];
            __syncwarp();
            const int m0 = (tile - mp.tile_start[p]) * F_BM + q * 32;
            for (int r = 0; r < 32; ++r) {
                const int row = m0 + r;
                if (row < M) {
                    float* y = Y + (long long)row * ldy;
                    y[lane] = st[r * 65 + lane] + bj[0];
                    y[32 + lane] = st[r * 65 + 32 + lane] + bj[1];
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, F_TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// forward, 3xTF32 ("x3"): fp32-class accuracy on the tensor cores.
//   A = A_hi + A_lo, W = W_hi + W_lo (each part exactly representable in TF32),  A W^T ~= A_hi W_hi + A_hi W_lo + A_lo W_hi
// W_hi / W_lo are pre-split in global memory (tiny); the A tile arrives raw by TMA and is split IN SHARED MEMORY by
// the four epilogue warps (element-wise, so the 128-byte swizzle is untouched: hi overwrites the raw tile, lo goes to a
// twin buffer at the same offset), fenced to the async proxy, then three MMAs per k-step read it.  Still HBM-bound:
// per 16 KB A tile the split costs ~300 SM cycles and the 12 MMAs ~200, the TMA stream ~700.
// ---------------------------------------------------------------------------------------------
constexpr int X_STAGES = 4;
constexpr int X_STAGE = 2 * F_A_BYTES + 2 * F_B_BYTES;   // A_hi | A_lo | W_hi | W_lo
constexpr int X_SMEM = X_STAGES * X_STAGE + 1024;

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// N float4 per thread (stride `nt` threads): every load is issued before the first store - hi and lo may alias as far as the
// compiler knows, so a load-convert-store loop is one dependent shared-memory round trip per element
template <int N>
__device__ __forceinline__ void split_f4(float4* hi, float4* lo, int t, int nt, int n_valid) {
    float4 x[N];
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (i < n_valid) x[i] = hi[i * nt + t];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (i < n_valid) {
            float4 h, l;
            h.x = tf32_rna(x[i].x); h.y = tf32_rna(x[i].y); h.z = tf32_rna(x[i].z); h.w = tf32_rna(x[i].w);
            l.x = tf32_rna(x[i].x - h.x); l.y = tf32_rna(x[i].y - h.y); l.z = tf32_rna(x[i].z - h.z); l.w = tf32_rna(x[i].w - h.w);
            hi[i * nt + t] = h;
            lo[i * nt + t] = l;
        }
    }
}


__global__ void __launch_bounds__(192)
linear_x3_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                     const __grid_constant__ CUtensorMap tmBl, const float* __restrict__ bias, float* __restrict__ Y,
                     long long ldy, int M, int num_kb) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[X_STAGES];
    __shared__ __align__(8) uint64_t split_bar[X_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[X_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * F_BM;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmBh);
        tma_prefetch_desc(&tmBl);
        for (int s = 0; s < X_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&split_bar[s], 128);   // every thread of the 4 split/epilogue warps arrives
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, F_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;

    if (warp == 0) {
        {   // whole warp in step, one elected lane issues (elect_one)
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % X_STAGES;
                const uint32_t ph = (kb / X_STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], F_A_BYTES + 2 * F_B_BYTES);
                uint8_t* a = smem + s * X_STAGE;
                tma_load_2d(&tmA, &full_bar[s], a, kb * F_BK, m0);
                tma_load_2d(&tmBh, &full_bar[s], a + 2 * F_A_BYTES, kb * F_BK, 0);
                tma_load_2d(&tmBl, &full_bar[s], a + 2 * F_A_BYTES + F_B_BYTES, kb * F_BK, 0);
            }
        }
    } else if (warp == 1) {
        {   // whole warp in step, one elected lane issues (elect_one)
            constexpr uint32_t idesc = umma_idesc_tf32(F_BM, F_BN, 0, 0);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % X_STAGES;
                const uint32_t ph = (kb / X_STAGES) & 1;
                mbar_wait(&split_bar[s], ph);
                tc_fence_after();
                const uint32_t ah = smem_u32(smem + s * X_STAGE);
                const uint32_t al = ah + F_A_BYTES;
                const uint32_t bh = ah + 2 * F_A_BYTES;
                const uint32_t bl = bh + F_B_BYTES;
#pragma unroll
                for (int k = 0; k < F_BK / 8; ++k) {
                    const uint64_t dah = umma_desc_sw128(ah + k * 32, 16, 1024);
                    const uint64_t dal = umma_desc_sw128(al + k * 32, 16, 1024);
                    const uint64_t dbh = umma_desc_sw128(bh + k * 32, 16, 1024);
                    const uint64_t dbl = umma_desc_sw128(bl + k * 32, 16, 1024);
                    umma_tf32(tmem_d, dal, dbh, idesc, (kb | k) != 0);   // small terms first
                    umma_tf32(tmem_d, dah, dbl, idesc, 1);
                    umma_tf32(tmem_d, dah, dbh, idesc, 1);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&tmem_full_bar);
        }
    } else {
        // ---- split phase: raw fp32 tile -> (hi, lo) TF32 pair, in place / twin buffer --------------------------------
        const int t = threadIdx.x - 64;   // 0..127
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % X_STAGES;
            const uint32_t ph = (kb / X_STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            float4* hi = reinterpret_cast<float4*>(smem + s * X_STAGE);
            float4* lo = reinterpret_cast<float4*>(smem + s * X_STAGE + F_A_BYTES);
            split_f4<F_A_BYTES / 16 / 128>(hi, lo, t, 128, F_A_BYTES / 16 / 128);
            fence_proxy_async();           // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(&split_bar[s]);
        }
        // ---- epilogue ---------------------------------------------------------------------------------------------
        const int q = warp & 3;
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        float v[64];
        const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
        tmem_ld_32x32(taddr, v);
        tmem_ld_32x32(taddr + 32, v + 32);
        tmem_ld_wait();
        float* st = reinterpret_cast<float*>(smem) + q * (32 * 65);
#pragma unroll
        for (int j = 0; j < 64; ++j) st[lane * 65 + j] = v[j] + (bias != nullptr ? __ldg(bias + j) : 0.f);
        __syncwarp();
        for (int r = 0; r < 32; ++r) {
            const int row = m0 + q * 32 + r;
            if (row < M) {
                float* y = Y + (long long)row * ldy;
                y[lane] = st[r * 65 + lane];
                y[32 + lane] = st[r * 65 + 32 + lane];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, F_TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// fused consumer of the layer-mean slab O [rows x 64(1+M)]  (3xTF32, one pass over O):
//   F[r]   = O[r, :]            @ Wf^T   + bf          (embedding_{user|item}_after_GCN, models/EliMRec.py:261-270)
//   S_m[r] = O[r, 64(m+1)..+64] @ Ws_m^T + bs_m        (s_dense_m, models/EliMRec.py:146-151)
// k-block kb covers O columns [32 kb, 32 kb + 32): it always feeds the fusion accumulator and, for kb >= 2, the head of
// graph block kb/2.  (1+M) accumulators of 64 columns live in TMEM; each A tile is split hi/lo in shared memory once and
// used by up to 6 MMAs per k-step.
// ---------------------------------------------------------------------------------------------
constexpr int H_STAGES = 3;
constexpr int H_STAGE = 2 * F_A_BYTES + 4 * F_B_BYTES;   // A_hi | A_lo | Wf_hi | Wf_lo | Ws_hi | Ws_lo
constexpr int H_SMEM = H_STAGES * H_STAGE + 1024;
constexpr int HA_SMEM = H_SMEM + 4 * 32 * 65 * 4;        // + one 32x64 transpose tile per drain warp (persistent variant)

struct HeadMaps {
    CUtensorMap hi[ELIMREC_MAX_MODS];
    CUtensorMap lo[ELIMREC_MAX_MODS];
};
struct HeadOut {
    const float* bias[1 + ELIMREC_MAX_MODS];
    float* out[1 + ELIMREC_MAX_MODS];      // each [rows x 64], row stride 64
};

__global__ void __launch_bounds__(192)
fuse_heads_x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmFh,
                     const __grid_constant__ CUtensorMap tmFl, const __grid_constant__ HeadMaps hm, HeadOut ho, int M_rows,
                     int n_heads) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[H_STAGES];
    __shared__ __align__(8) uint64_t split_bar[H_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[H_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * F_BM;
    const int num_kb = 2 * (1 + n_heads);
    const uint32_t tmem_cols = (n_heads >= 2) ? 256u : 128u;   // power of two >= 64 (1 + n_heads)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmFh);
        tma_prefetch_desc(&tmFl);
        for (int s = 0; s < H_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&split_bar[s], 128);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;

    if (warp == 0) {
        {   // whole warp in step, one elected lane issues (elect_one)
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % H_STAGES;
                const uint32_t ph = (kb / H_STAGES) & 1;
                const int head = kb / 2 - 1;   // -1: identity block, no head
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], F_A_BYTES + 2 * F_B_BYTES + (head >= 0 ? 2 * F_B_BYTES : 0));
                uint8_t* a = smem + s * H_STAGE;
                uint8_t* b = a + 2 * F_A_BYTES;
                tma_load_2d(&tmA, &full_bar[s], a, kb * F_BK, m0);
                tma_load_2d(&tmFh, &full_bar[s], b, kb * F_BK, 0);
                tma_load_2d(&tmFl, &full_bar[s], b + F_B_BYTES, kb * F_BK, 0);
                if (head >= 0) {
                    tma_load_2d(&hm.hi[head], &full_bar[s], b + 2 * F_B_BYTES, (kb & 1) * F_BK, 0);
                    tma_load_2d(&hm.lo[head], &full_bar[s], b + 3 * F_B_BYTES, (kb & 1) * F_BK, 0);
                }
            }
        }
    } else if (warp == 1) {
        {   // whole warp in step, one elected lane issues (elect_one)
            constexpr uint32_t idesc = umma_idesc_tf32(F_BM, F_BN, 0, 0);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % H_STAGES;
                const uint32_t ph = (kb / H_STAGES) & 1;
                const int head = kb / 2 - 1;
                mbar_wait(&split_bar[s], ph);
                tc_fence_after();
                const uint32_t ah = smem_u32(smem + s * H_STAGE);
                const uint32_t al = ah + F_A_BYTES;
                const uint32_t fh = ah + 2 * F_A_BYTES, fl = fh + F_B_BYTES, sh = fl + F_B_BYTES, sl = sh + F_B_BYTES;
#pragma unroll
                for (int k = 0; k < F_BK / 8; ++k) {
                    const uint64_t dah = umma_desc_sw128(ah + k * 32, 16, 1024);
                    const uint64_t dal = umma_desc_sw128(al + k * 32, 16, 1024);
                    const uint64_t dfh = umma_desc_sw128(fh + k * 32, 16, 1024);
                    const uint64_t dfl = umma_desc_sw128(fl + k * 32, 16, 1024);
                    umma_tf32(tmem_d, dal, dfh, idesc, (kb | k) != 0);
                    umma_tf32(tmem_d, dah, dfl, idesc, 1);
                    umma_tf32(tmem_d, dah, dfh, idesc, 1);
                    if (head >= 0) {
                        const uint64_t dsh = umma_desc_sw128(sh + k * 32, 16, 1024);
                        const uint64_t dsl = umma_desc_sw128(sl + k * 32, 16, 1024);
                        const uint32_t dcol = tmem_d + 64 * (head + 1);
                        umma_tf32(dcol, dal, dsh, idesc, ((kb & 1) | k) != 0);
                        umma_tf32(dcol, dah, dsl, idesc, 1);
                        umma_tf32(dcol, dah, dsh, idesc, 1);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&tmem_full_bar);
        }
    } else {
        const int t = threadIdx.x - 64;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % H_STAGES;
            const uint32_t ph = (kb / H_STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            float4* hi = reinterpret_cast<float4*>(smem + s * H_STAGE);
            float4* lo = reinterpret_cast<float4*>(smem + s * H_STAGE + F_A_BYTES);
            split_f4<F_A_BYTES / 16 / 128>(hi, lo, t, 128, F_A_BYTES / 16 / 128);
            fence_proxy_async();
            mbar_arrive(&split_bar[s]);
        }
        const int q = warp & 3;
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        float* st = reinterpret_cast<float*>(smem) + q * (32 * 65);
        for (int o = 0; o <= n_heads; ++o) {
            float v[64];
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + 64 * o;
            tmem_ld_32x32(taddr, v);
            tmem_ld_32x32(taddr + 32, v + 32);
            tmem_ld_wait();
            const float* bias = ho.bias[o];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 64; ++j) st[lane * 65 + j] = v[j] + (bias != nullptr ? __ldg(bias + j) : 0.f);
            __syncwarp();
            float* outp = ho.out[o];
            for (int r = 0; r < 32; ++r) {
                const int row = m0 + q * 32 + r;
                if (row < M_rows) {
                    float* y = outp + (long long)row * 64;
                    y[lane] = st[r * 65 + lane];
                    y[32 + lane] = st[r * 65 + 32 + lane];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// Persistent variant over the WHOLE slab O [N x 64(1+M)] (users first, then items): one CTA per SM walks 128-row tiles
// (user tiles use the user fusion weights, item tiles the item ones), the accumulators are double-buffered in TMEM
// (2 x 256 columns) and separate warps split (4), issue MMAs (1), load (1) and drain (4), so the epilogue of tile i
// overlaps the main loop of tile i+1 and the prologue is paid once per SM instead of once per tile.
// ---------------------------------------------------------------------------------------------
struct FuseAllMaps {
    CUtensorMap A;            // O, all N rows
    CUtensorMap Fh[2], Fl[2]; // fusion weights hi/lo: [0] users, [1] items
    CUtensorMap Sh[ELIMREC_MAX_MODS], Sl[ELIMREC_MAX_MODS];
};
struct FuseAllOut {
    const float* bias_f[2];
    const float* bias_s[ELIMREC_MAX_MODS];
    float* out[1 + ELIMREC_MAX_MODS];   // [N x 64] each: fused table, then the heads
};

__global__ void __launch_bounds__(320, 1)
fuse_heads_x3_all_kernel(const __grid_constant__ FuseAllMaps mp, FuseAllOut ho, int U, int N, int n_heads, int n_user_tiles,
                         int n_tiles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[H_STAGES];
    __shared__ __align__(8) uint64_t split_bar[H_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[H_STAGES];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = 2 * (1 + n_heads);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mp.A);
        for (int s = 0; s < H_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&split_bar[s], 128);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], 4);   // one arrival per drain warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;

    if (warp == 0) {
        {   // whole warp in step, one elected lane issues (elect_one)
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int who = tile < n_user_tiles ? 0 : 1;
                const int row0 = who == 0 ? tile * F_BM : U + (tile - n_user_tiles) * F_BM;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % H_STAGES;
                    const uint32_t ph = (it / H_STAGES) & 1;
                    const int head = kb / 2 - 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_expect_tx(&full_bar[s], F_A_BYTES + 2 * F_B_BYTES + (head >= 0 ? 2 * F_B_BYTES : 0));
                    uint8_t* a = smem + s * H_STAGE;
                    uint8_t* b = a + 2 * F_A_BYTES;
                    tma_load_2d(&mp.A, &full_bar[s], a, kb * F_BK, row0);
                    tma_load_2d(&mp.Fh[who], &full_bar[s], b, kb * F_BK, 0);
                    tma_load_2d(&mp.Fl[who], &full_bar[s], b + F_B_BYTES, kb * F_BK, 0);
                    if (head >= 0) {
                        tma_load_2d(&mp.Sh[head], &full_bar[s], b + 2 * F_B_BYTES, (kb & 1) * F_BK, 0);
                        tma_load_2d(&mp.Sl[head], &full_bar[s], b + 3 * F_B_BYTES, (kb & 1) * F_BK, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        {   // whole warp in step, one elected lane issues (elect_one)
            constexpr uint32_t idesc = umma_idesc_tf32(F_BM, F_BN, 0, 0);
            int it = 0, tc = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tc) {
                const int buf = tc & 1;
                mbar_wait(&tempty_bar[buf], ((tc >> 1) & 1) ^ 1);   // drained by the epilogue two tiles ago
                tc_fence_after();
                const uint32_t dbase = tmem_d + buf * 256;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % H_STAGES;
                    const uint32_t ph = (it / H_STAGES) & 1;
                    const int head = kb / 2 - 1;
                    mbar_wait(&split_bar[s], ph);
                    tc_fence_after();
                    const uint32_t ah = smem_u32(smem + s * H_STAGE);
                    const uint32_t al = ah + F_A_BYTES;
                    const uint32_t fh = ah + 2 * F_A_BYTES, fl = fh + F_B_BYTES, sh = fl + F_B_BYTES, sl = sh + F_B_BYTES;
#pragma unroll
                    for (int k = 0; k < F_BK / 8; ++k) {
                        const uint64_t dah = umma_desc_sw128(ah + k * 32, 16, 1024);
                        const uint64_t dal = umma_desc_sw128(al + k * 32, 16, 1024);
                        const uint64_t dfh = umma_desc_sw128(fh + k * 32, 16, 1024);
                        const uint64_t dfl = umma_desc_sw128(fl + k * 32, 16, 1024);
                        umma_tf32(dbase, dal, dfh, idesc, (kb | k) != 0);
                        umma_tf32(dbase, dah, dfl, idesc, 1);
                        umma_tf32(dbase, dah, dfh, idesc, 1);
                        if (head >= 0) {
                            const uint64_t dsh = umma_desc_sw128(sh + k * 32, 16, 1024);
                            const uint64_t dsl = umma_desc_sw128(sl + k * 32, 16, 1024);
                            const uint32_t dcol = dbase + 64 * (head + 1);
                            umma_tf32(dcol, dal, dsh, idesc, ((kb & 1) | k) != 0);
                            umma_tf32(dcol, dah, dsl, idesc, 1);
                            umma_tf32(dcol, dah, dsh, idesc, 1);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tfull_bar[buf]);
            }
        }
    } else if (warp < 6) {
        // ---- split warps: raw fp32 A tile -> TF32 hi (in place) + lo (twin buffer) --------------------------------------
        const int t = threadIdx.x - 64;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const int s = it % H_STAGES;
                const uint32_t ph = (it / H_STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                float4* hi = reinterpret_cast<float4*>(smem + s * H_STAGE);
                float4* lo = reinterpret_cast<float4*>(smem + s * H_STAGE + F_A_BYTES);
                split_f4<F_A_BYTES / 16 / 128>(hi, lo, t, 128, F_A_BYTES / 16 / 128);
                fence_proxy_async();
                mbar_arrive(&split_bar[s]);
            }
        }
    } else {
        // ---- drain warps: TMEM -> registers -> global (each thread owns one row: 64 consecutive floats per output) -----
        const int q = warp & 3;
        int tc = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tc) {
            const int buf = tc & 1;
            const int who = tile < n_user_tiles ? 0 : 1;
            const int row0 = who == 0 ? tile * F_BM : U + (tile - n_user_tiles) * F_BM;
            const int row_end = who == 0 ? U : N;
            const int row = row0 + q * 32 + lane;
            mbar_wait(&tfull_bar[buf], (tc >> 1) & 1);
            tc_fence_after();
            for (int o = 0; o <= n_heads; ++o) {
                float v[64];
                const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + buf * 256 + 64 * o;
                tmem_ld_32x32(taddr, v);
                tmem_ld_32x32(taddr + 32, v + 32);
                tmem_ld_wait();
                if (o == n_heads) {   // last read of this buffer: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[buf]);
                }
                // transpose through a private shared-memory tile so that every store instruction writes whole rows
                // (a thread-per-row float4 store touches 32 different 128-byte lines per instruction: LSU-throttled)
                const float* bias = (o == 0) ? ho.bias_f[who] : ho.bias_s[o - 1];
                float* st = reinterpret_cast<float*>(smem + H_STAGES * H_STAGE) + q * (32 * 65);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 64; ++j) st[lane * 65 + j] = v[j];
                __syncwarp();
                const float b0 = __ldg(bias + lane), b1 = __ldg(bias + 32 + lane);
                float* outp = ho.out[o];
                const int rbase = row0 + q * 32;
                for (int r = 0; r < 32; ++r) {
                    if (rbase + r < row_end) {
                        float* y = outp + (long long)(rbase + r) * 64;
                        y[lane] = st[r * 65 + lane] + b0;
                        y[32 + lane] = st[r * 65 + 32 + lane] + b1;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, 512);
}

__global__ void split_tf32_kernel(long long n, const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = src[i];
    const float h = tf32_rna(x);
    hi[i] = h;
    lo[i] = tf32_rna(x - h);
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------
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

// row-major fp32 [rows x cols] with row stride ld (floats); box = [box_rows x 32 floats], 128-byte swizzle,
// out-of-bounds elements read as zero (handles the M and K tails)
int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
             CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) return -1;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------------
// weight gradient:  dW[n, c] = sum_m dY[m, n] * X[m, c]      (computed as D^T[c, n], both operands MN-major)
// ---------------------------------------------------------------------------------------------
constexpr int G_BR = 32;                 // rows of X / dY per pipeline stage (4 MMA k-steps of 8 rows)
constexpr int G_KC = 256;                // columns of X owned by one CTA (two 128-wide MMA tiles)
constexpr int G_STAGES = 4;
constexpr int G_BOX = G_BR * 128;        // one TMA box: [32 rows x 32 floats] = 4 KB
constexpr int G_XB = G_KC / 32;          // X boxes per stage
constexpr int G_STAGE = (G_XB + 2) * G_BOX;
constexpr int G_SMEM = G_STAGES * G_STAGE + 1024;
constexpr int G_TMEM_COLS = 128;

__global__ void __launch_bounds__(192)
linear_tf32_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                         float* __restrict__ part, int M, int K, int rb_per_cta) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[G_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[G_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * G_KC;                     // first X column of this CTA
    const int n_tiles = (min(K - c0, G_KC) + 127) / 128;  // 1 or 2 MMA tiles
    const int n_boxes = (min(K - c0, G_KC) + 31) / 32;    // X boxes that contain any valid column
    const int nrb = (M + G_BR - 1) / G_BR;
    const int rb0 = blockIdx.y * rb_per_cta;
    const int rb1 = min(nrb, rb0 + rb_per_cta);
    const int n_it = max(0, rb1 - rb0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmG);
        for (int s = 0; s < G_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, G_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;

    if (warp == 0) {
        {   // whole warp in step, one elected lane issues (elect_one)
            for (int it = 0; it < n_it; ++it) {
                const int s = it % G_STAGES;
                const uint32_t ph = (it / G_STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], (uint32_t)(n_boxes + 2) * G_BOX);
                uint8_t* base = smem + s * G_STAGE;
                const int row = (rb0 + it) * G_BR;
                for (int b = 0; b < n_boxes; ++b) tma_load_2d(&tmX, &full_bar[s], base + b * G_BOX, c0 + b * 32, row);
                tma_load_2d(&tmG, &full_bar[s], base + G_XB * G_BOX, 0, row);
                tma_load_2d(&tmG, &full_bar[s], base + (G_XB + 1) * G_BOX, 32, row);
            }
        }
    } else if (warp == 1) {
        {   // whole warp in step, one elected lane issues (elect_one)
            constexpr uint32_t idesc = umma_idesc_tf32(128, 64, 1, 1);  // A = X^T and B = dY are both MN-major
            for (int it = 0; it < n_it; ++it) {
                const int s = it % G_STAGES;
                const uint32_t ph = (it / G_STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t xb = smem_u32(smem + s * G_STAGE);
                const uint32_t gb = xb + G_XB * G_BOX;
                for (int t = 0; t < n_tiles; ++t) {
#pragma unroll
                    for (int j = 0; j < G_BR / 8; ++j) {  // one MMA = 8 rows = two 4-row (512-byte) swizzle atoms per 32-column group
                        const uint64_t da = umma_desc(xb + t * 4 * G_BOX + j * 1024, G_BOX, 512, 1);
                        const uint64_t db = umma_desc(gb + j * 1024, G_BOX, 512, 1);
                        umma_tf32(tmem_d + t * 64, da, db, idesc, (it | j) != 0);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&tmem_full_bar);
        }
    } else {
        const int q = warp & 3;
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        float* out = part + (long long)blockIdx.y * 64 * K;   // this row-range's partial, laid out like dW [64 x K]
        for (int t = 0; t < n_tiles; ++t) {
            float v[64];
            const int c = c0 + t * 128 + q * 32 + lane;
            if (n_it > 0) {
                const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + t * 64;
                tmem_ld_32x32(taddr, v);
                tmem_ld_32x32(taddr + 32, v + 32);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int n = 0; n < 64; ++n) v[n] = 0.f;
            }
            if (c < K) {
#pragma unroll
                for (int n = 0; n < 64; ++n) out[(long long)n * K + c] = v[n];  // 32 consecutive c per warp store
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, G_TMEM_COLS);
}

__global__ void wgrad_reduce_kernel(long long n, int n_part, const float* __restrict__ part, float* __restrict__ dW) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int p = 0; p < n_part; ++p) s += part[(long long)p * n + i];  // fixed order: deterministic
    dW[i] = s;
}

void wgrad_plan(int64_t M, int64_t K, int* nk, int* rs, int* rb_per_cta) {
    *nk = (int)((K + G_KC - 1) / G_KC);
    const int nrb = (int)((M + G_BR - 1) / G_BR);
    int r = 148 / *nk;
    if (r < 1) r = 1;
    if (r > nrb) r = nrb;
    *rb_per_cta = (nrb + r - 1) / r;
    *rs = (nrb + *rb_per_cta - 1) / *rb_per_cta;
}

}  // namespace

ELIMREC_API int64_t elimrec_linear_tf32_wgrad_workspace_floats(int64_t M, int64_t K) {
    int nk, rs, rbp;
    wgrad_plan(M, K, &nk, &rs, &rbp);
    return (int64_t)rs * 64 * K;
}

ELIMREC_API int elimrec_linear_tf32_wgrad(int64_t M, int64_t K, const float* dY, int64_t lddy, const float* X, int64_t ldx,
                                          float* dW, float* workspace, elimrec_stream_t stream) {
    if (M <= 0 || K <= 0) return 0;
    if (K % 4 != 0 || ldx % 4 != 0 || lddy % 4 != 0 || !aligned16(X) || !aligned16(dY) || M > 0x7fffffff || workspace == nullptr) {
        elimrec_set_error("elimrec_linear_tf32_wgrad: unsupported shape/alignment (K=%lld ldx=%lld lddy=%lld)", (long long)K,
                          (long long)ldx, (long long)lddy);
        return -2;
    }
    CUtensorMap tmX, tmG;
    if (make_map(&tmX, X, M, K, ldx, G_BR, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) != 0 ||
        make_map(&tmG, dY, M, 64, lddy, G_BR, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) != 0) {
        elimrec_set_error("elimrec_linear_tf32_wgrad: cuTensorMapEncodeTiled failed");
        return -3;
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(linear_tf32_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
        if (e != cudaSuccess) {
            elimrec_set_error("elimrec_linear_tf32_wgrad: shared-memory opt-in failed: %s", cudaGetErrorString(e));
            return -3;
        }
        configured = true;
    }
    int nk, rs, rbp;
    wgrad_plan(M, K, &nk, &rs, &rbp);
    cudaStream_t st = er_stream(stream);
    linear_tf32_wgrad_kernel<<<dim3(nk, rs), 192, G_SMEM, st>>>(tmX, tmG, workspace, (int)M, (int)K, rbp);
    ER_LAUNCH_CHECK();
    const long long n = 64LL * K;
    wgrad_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, rs, workspace, dW);
    ER_LAUNCH_CHECK();
    return 0;
}

namespace {
__global__ void round_tf32_kernel(long long n, const float* __restrict__ src, float* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(src[i]));  // round-to-nearest (the MMA itself truncates)
    dst[i] = __uint_as_float(r);
}
}  // namespace

ELIMREC_API int elimrec_round_tf32(int64_t n, const float* src, float* dst, elimrec_stream_t stream) {
    if (n <= 0) return 0;
    round_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, er_stream(stream)>>>(n, src, dst);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_linear_tf32_fwd_multi(int n, const elimrec_linear_desc_t* d, elimrec_stream_t stream) {
    ER_CHECK_ARG(n >= 0 && n <= F_MAX_PROB && (n == 0 || d != nullptr), "at most 4 problems per launch");
    FwdMulti mp{};
    int order[F_MAX_PROB], cnt = 0;
    for (int i = 0; i < n; ++i)
        if (d[i].M > 0) order[cnt++] = i;
    for (int i = 1; i < cnt; ++i)            // largest K first: the long tiles are dealt before the short ones
        for (int j = i; j > 0 && d[order[j]].K > d[order[j - 1]].K; --j) { int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
    if (cnt == 0) return 0;
    for (int i = 0; i < cnt; ++i) {
        const elimrec_linear_desc_t& q = d[order[i]];
        if (q.K < 4 || q.K % 4 != 0 || q.ldx % 4 != 0 || !aligned16(q.X) || !aligned16(q.W) || q.M > 0x7fffffff) {
            elimrec_set_error("elimrec_linear_tf32_fwd: unsupported shape/alignment (K=%lld ldx=%lld)", (long long)q.K, (long long)q.ldx);
            return -2;
        }
        if (make_map(&mp.A[i], q.X, q.M, q.K, q.ldx, F_BM) != 0 || make_map(&mp.B[i], q.W, 64, q.K, q.K, F_BN) != 0) {
            elimrec_set_error("elimrec_linear_tf32_fwd: cuTensorMapEncodeTiled failed");
            return -3;
        }
        mp.bias[i] = q.b; mp.Y[i] = q.Y; mp.ldy[i] = q.ldy; mp.M[i] = (int)q.M;
        mp.num_kb[i] = (int)((q.K + F_BK - 1) / F_BK);
        mp.tile_start[i + 1] = mp.tile_start[i] + (int)((q.M + F_BM - 1) / F_BM);
    }
    mp.n = cnt;
    static bool configured = false;
    static int n_sm = 148;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(linear_tf32_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM);
        if (e != cudaSuccess) {
            elimrec_set_error("elimrec_linear_tf32_fwd: shared-memory opt-in failed: %s", cudaGetErrorString(e));
            return -3;
        }
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    const int n_tiles = mp.tile_start[cnt];
    const unsigned grid = (unsigned)(n_tiles < n_sm ? n_tiles : n_sm);
    linear_tf32_fwd_kernel<<<grid, 192, F_SMEM, er_stream(stream)>>>(mp);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_linear_tf32_fwd(int64_t M, int64_t K, const float* X, int64_t ldx, const float* W, const float* b,
                                        float* Y, int64_t ldy, elimrec_stream_t stream) {
    elimrec_linear_desc_t d{M, K, X, ldx, W, b, Y, ldy};
    return elimrec_linear_tf32_fwd_multi(1, &d, stream);
}




// ---------------------------------------------------------------------------------------------
// Batched 3xTF32 weight gradients (the tensor-core twin of csrc/wgrad_multi.cu, same problems / workspace / reduction):
//   partial[tile][split][n][c] = sum over the split's rows of  A_p[m, n] * B_p[m, c0 + c]     (+ column sums of A_p)
// computed like linear_tf32_wgrad_kernel as D^T[c, n] with both operands MN-major straight from row-major memory, but on
// fp32-accurate operands: four split warps turn every landed TMA stage into TF32 hi (in place) + lo (twin buffer) and the
// MMA thread issues lo*hi + hi*lo + hi*hi into one TMEM accumulator.  The bias gradient (column sums of A) rides along as
// two more MMAs against a constant all-ones A' tile into a second 64-column accumulator.
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int X3_BR = 32;                        // rows per stage (4 MMA k-steps of 8 rows)
constexpr int X3_STAGES = 4;
constexpr int X3_BOX = X3_BR * 128;              // one TMA box: [32 rows x 32 floats] = 4 KB
constexpr int X3_G = 2 * X3_BOX;                 // A_p block: 64 columns
constexpr int X3_X = 4 * X3_BOX;                 // B_p tile: 128 columns
constexpr int X3_STAGE = 2 * X3_G + 2 * X3_X;    // [G hi | G lo | X hi | X lo] = 48 KB
constexpr int X3_ONES = 4096;
constexpr int X3_SMEM = X3_STAGES * X3_STAGE + X3_ONES + 1024;
constexpr int X3_TMEM_COLS = 512;              // 4 rotating [128 x 64] accumulators + 4 for the bias sums
constexpr int X3_SPLIT = 256;                    // split / epilogue threads (warps 2..9)
constexpr int X3_THREADS = 64 + X3_SPLIT;

struct X3Args {
    CUtensorMap A[ELIMREC_WGRAD_MAX_PROBLEMS], B[ELIMREC_WGRAD_MAX_PROBLEMS];
    int K[ELIMREC_WGRAD_MAX_PROBLEMS], rows[ELIMREC_WGRAD_MAX_PROBLEMS], has_bias[ELIMREC_WGRAD_MAX_PROBLEMS];
    int tile0[ELIMREC_WGRAD_MAX_PROBLEMS + 1];
    int n_prob, n_tiles, splits;
    float* ws;
};

__global__ void __launch_bounds__(X3_THREADS, 3) wgrad_x3_multi_kernel(const __grid_constant__ X3Args a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[X3_STAGES];
    __shared__ __align__(8) uint64_t split_bar[X3_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[X3_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, split = blockIdx.y;
    int p = 0;
    while (p + 1 < a.n_prob && tile >= a.tile0[p + 1]) ++p;
    const int c0 = (tile - a.tile0[p]) * 128;              // first B column of this tile
    const int K = a.K[p], rows = a.rows[p];
    const int n_boxes = (min(K - c0, 128) + 31) / 32;      // B boxes holding any valid column
    int chunk = (rows + a.splits - 1) / a.splits;
    chunk = (chunk + X3_BR - 1) / X3_BR * X3_BR;
    const int rb0 = split * chunk;
    const int rb1 = min(rows, rb0 + chunk);
    const int n_it = rb1 > rb0 ? (rb1 - rb0 + X3_BR - 1) / X3_BR : 0;    // the last stage's rows past `rows` arrive as zeros
    const bool want_bias = a.has_bias[p] != 0 && c0 == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&a.A[p]);
        tma_prefetch_desc(&a.B[p]);
        for (int s = 0; s < X3_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&split_bar[s], X3_SPLIT);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, X3_TMEM_COLS);
    if (warp >= 2 && want_bias) {                          // the all-ones A' tile of the bias accumulator
        float4* ones = reinterpret_cast<float4*>(smem + X3_STAGES * X3_STAGE);
        for (int i = threadIdx.x - 64; i < X3_ONES / 16; i += X3_SPLIT) ones[i] = make_float4(1.f, 1.f, 1.f, 1.f);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;

    if (warp == 0) {
        {   // whole warp in step, one elected lane issues (elect_one)
            for (int it = 0; it < n_it; ++it) {
                const int s = it % X3_STAGES;
                const uint32_t ph = (it / X3_STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], (uint32_t)(n_boxes + 2) * X3_BOX);
                uint8_t* base = smem + s * X3_STAGE;
                const int row = rb0 + it * X3_BR;
                tma_load_2d(&a.A[p], &full_bar[s], base, 0, row);
                tma_load_2d(&a.A[p], &full_bar[s], base + X3_BOX, 32, row);
                for (int b = 0; b < n_boxes; ++b) tma_load_2d(&a.B[p], &full_bar[s], base + 2 * X3_G + b * X3_BOX, c0 + b * 32, row);
            }
        }
    } else if (warp == 1) {
        {   // whole warp in step, one elected lane issues (elect_one)
            constexpr uint32_t idesc = umma_idesc_tf32(128, 64, 1, 1);   // A' = B_p^T and B' = A_p are both MN-major
            const uint32_t ones = smem_u32(smem + X3_STAGES * X3_STAGE);
            for (int it = 0; it < n_it; ++it) {
                const int s = it % X3_STAGES;
                const uint32_t ph = (it / X3_STAGES) & 1;
                mbar_wait(&split_bar[s], ph);
                tc_fence_after();
                const uint32_t gh = smem_u32(smem + s * X3_STAGE), gl = gh + X3_G;
                const uint32_t xh = gl + X3_G, xl = xh + X3_X;
#pragma unroll
                for (int j = 0; j < X3_BR / 8; ++j) {      // one MMA = 8 rows = two 4-row (512-byte) swizzle atoms per 32-column group
                    const uint64_t dxh = umma_desc(xh + j * 1024, X3_BOX, 512, 1), dxl = umma_desc(xl + j * 1024, X3_BOX, 512, 1);
                    const uint64_t dgh = umma_desc(gh + j * 1024, X3_BOX, 512, 1), dgl = umma_desc(gl + j * 1024, X3_BOX, 512, 1);
                    // k-step j of every stage goes to accumulator j: the tensor core's fp32 accumulate is not a rounded fp32 add
                    // (measured: ~2^-24 of the running sum lost per MMA, same sign), so four shorter chains summed once in
                    // registers carry a quarter of the error of one long chain
                    const uint32_t dj = tmem_d + j * 64;
                    umma_tf32(dj, dxl, dgh, idesc, it != 0);
                    umma_tf32(dj, dxh, dgl, idesc, 1);
                    umma_tf32(dj, dxh, dgh, idesc, 1);
                    if (want_bias) {
                        const uint64_t d1 = umma_desc(ones, 1024, 512, 1);
                        umma_tf32(dj + 256, d1, dgl, idesc, it != 0);
                        umma_tf32(dj + 256, d1, dgh, idesc, 1);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&tmem_full_bar);
        }
    } else {
        const int t = threadIdx.x - 64;
        for (int it = 0; it < n_it; ++it) {
            const int s = it % X3_STAGES;
            const uint32_t ph = (it / X3_STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            uint8_t* base = smem + s * X3_STAGE;
            float4* ghi = reinterpret_cast<float4*>(base);
            float4* glo = reinterpret_cast<float4*>(base + X3_G);
            float4* xhi = reinterpret_cast<float4*>(base + 2 * X3_G);
            float4* xlo = reinterpret_cast<float4*>(base + 2 * X3_G + X3_X);
            split_f4<X3_G / 16 / X3_SPLIT>(ghi, glo, t, X3_SPLIT, X3_G / 16 / X3_SPLIT);
            split_f4<X3_X / 16 / X3_SPLIT>(xhi, xlo, t, X3_SPLIT, n_boxes * (X3_BOX / 16 / X3_SPLIT));
            fence_proxy_async();
            mbar_arrive(&split_bar[s]);
        }
        const int q = warp & 3, half = (warp - 2) >> 2;        // TMEM lane quadrant of this warp, and which 32 of the 64 columns it drains
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        float* out = a.ws + ((long long)tile * a.splits + split) * (64 * 128);     // [64 x 128] like wgrad_multi's partial tiles
        // eight columns at a time: the kernel runs beside the propagation launches, which want the register file - a lean
        // drain keeps this CTA at 64 registers per thread so that two propagation CTAs fit on the SM next to it
        auto drain4 = [&](uint32_t taddr, float* dst, int stride) {   // (acc0 + acc1) + (acc2 + acc3) of the rotating accumulators
#pragma unroll 1
            for (int c = 0; c < 32; c += 8) {
                float v[8], w[8];
                if (n_it > 0) {
                    tmem_ld_32x8(taddr + c, v);
                    tmem_ld_32x8(taddr + c + 64, w);
                    tmem_ld_wait();
#pragma unroll
                    for (int n = 0; n < 8; ++n) v[n] += w[n];
                    float x[8];
                    tmem_ld_32x8(taddr + c + 128, w);
                    tmem_ld_32x8(taddr + c + 192, x);
                    tmem_ld_wait();
#pragma unroll
                    for (int n = 0; n < 8; ++n) v[n] += w[n] + x[n];
                } else {
#pragma unroll
                    for (int n = 0; n < 8; ++n) v[n] = 0.f;
                }
                if (dst != nullptr) {
#pragma unroll
                    for (int n = 0; n < 8; ++n) dst[(c + n) * stride] = v[n];
                }
            }
        };
        // partial tile [64 x 128] like wgrad_multi's: 32 consecutive c per warp store
        drain4(tmem_d + ((uint32_t)(q * 32) << 16) + half * 32, out + (half * 32) * 128 + q * 32 + lane, 128);
        if (want_bias && q == 0) {
            float* bp = a.ws + (long long)a.n_tiles * a.splits * (64 * 128) + ((long long)tile * a.splits + split) * 64 + half * 32;
            drain4(tmem_d + 256 + half * 32, lane == 0 ? bp : nullptr, 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, X3_TMEM_COLS);
}
}  // namespace

int er_wgrad_multi_reduce(int n, const elimrec_wgrad_problem_t* problems, int splits, float* workspace, const float* gscale_dev,
                          cudaStream_t st);   // csrc/wgrad_multi.cu

ELIMREC_API int elimrec_wgrad_multi_x3(int n, const elimrec_wgrad_problem_t* problems, int splits, float* workspace,
                                       const float* gscale_dev, elimrec_stream_t stream) {
    ER_CHECK_ARG(n >= 0 && n <= ELIMREC_WGRAD_MAX_PROBLEMS && (n == 0 || problems != nullptr), "too many problems");
    ER_CHECK_ARG(splits >= 1 && splits <= 64 && workspace != nullptr, "splits in [1, 64] and a workspace required");
    if (n == 0) return 0;
    X3Args a{};
    int tiles = 0;
    for (int i = 0; i < n; ++i) {
        const elimrec_wgrad_problem_t& q = problems[i];
        ER_CHECK_ARG(q.K > 0 && q.row_begin >= 0 && q.row_end >= q.row_begin && q.row_end <= 0x7fffffff, "bad problem shape");
        const float* A = q.A + q.row_begin * q.lda;
        const float* B = q.B + q.row_begin * q.ldb;
        if (q.lda % 4 != 0 || q.ldb % 4 != 0 || !aligned16(A) || !aligned16(B)) {
            elimrec_set_error("elimrec_wgrad_multi_x3: operands must be 16-byte aligned with row strides that are multiples of 4 "
                              "(problem %d: lda=%lld ldb=%lld)", i, (long long)q.lda, (long long)q.ldb);
            return -2;
        }
        const int64_t rows = q.row_end - q.row_begin;
        if (rows > 0 && (make_map(&a.A[i], A, rows, 64, q.lda, X3_BR, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) != 0 ||
                         make_map(&a.B[i], B, rows, q.K, q.ldb, X3_BR, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) != 0)) {
            elimrec_set_error("elimrec_wgrad_multi_x3: cuTensorMapEncodeTiled failed (problem %d)", i);
            return -3;
        }
        a.K[i] = (int)q.K;
        a.rows[i] = (int)rows;
        a.has_bias[i] = q.bias_out != nullptr;
        a.tile0[i] = tiles;
        tiles += (int)((q.K + 127) / 128);
    }
    a.tile0[n] = tiles;
    a.n_prob = n;
    a.n_tiles = tiles;
    a.splits = splits;
    a.ws = workspace;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_x3_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, X3_SMEM);
        if (e != cudaSuccess) {
            elimrec_set_error("elimrec_wgrad_multi_x3: shared-memory opt-in failed: %s", cudaGetErrorString(e));
            return -3;
        }
        configured = true;
    }
    cudaStream_t st = er_stream(stream);
    wgrad_x3_multi_kernel<<<dim3(tiles, splits), X3_THREADS, X3_SMEM, st>>>(a);
    ER_LAUNCH_CHECK();
    return er_wgrad_multi_reduce(n, problems, splits, workspace, gscale_dev, st);
}

namespace {
struct PrepMulti {
    int n;
    const float* src[ELIMREC_PREP_MAX];
    float* hi[ELIMREC_PREP_MAX];
    float* lo[ELIMREC_PREP_MAX];   // NULL: round only
    long long numel[ELIMREC_PREP_MAX];
    int block_start[ELIMREC_PREP_MAX + 1];
};
__global__ void prep_multi_kernel(const __grid_constant__ PrepMulti a) {
    int t = 0;
    while (t + 1 < a.n && (int)blockIdx.x >= a.block_start[t + 1]) ++t;
    const long long i = (long long)(blockIdx.x - a.block_start[t]) * 256 + threadIdx.x;
    if (i >= a.numel[t]) return;
    const float x = a.src[t][i];
    const float h = tf32_rna(x);
    a.hi[t][i] = h;
    if (a.lo[t] != nullptr) a.lo[t][i] = tf32_rna(x - h);
}
}  // namespace

ELIMREC_API int elimrec_prep_weights_tf32(int n, const elimrec_prep_tensor_t* t, elimrec_stream_t stream) {
    ER_CHECK_ARG(n >= 0 && n <= ELIMREC_PREP_MAX, "too many tensors");
    if (n == 0) return 0;
    PrepMulti a{};
    a.n = n;
    int blocks = 0;
    for (int i = 0; i < n; ++i) {
        a.src[i] = t[i].src; a.hi[i] = t[i].hi; a.lo[i] = t[i].lo; a.numel[i] = t[i].numel;
        a.block_start[i] = blocks;
        blocks += (int)((t[i].numel + 255) / 256);
    }
    a.block_start[n] = blocks;
    if (blocks == 0) return 0;
    prep_multi_kernel<<<blocks, 256, 0, er_stream(stream)>>>(a);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_split_tf32(int64_t n, const float* src, float* hi, float* lo, elimrec_stream_t stream) {
    if (n <= 0) return 0;
    split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, er_stream(stream)>>>(n, src, hi, lo);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_linear_x3_fwd(int64_t M, int64_t K, const float* X, int64_t ldx, const float* W_hi, const float* W_lo,
                                      const float* b, float* Y, int64_t ldy, elimrec_stream_t stream) {
    if (M <= 0) return 0;
    if (K < 4 || K % 4 != 0 || ldx % 4 != 0 || !aligned16(X) || !aligned16(W_hi) || !aligned16(W_lo) || M > 0x7fffffff) {
        elimrec_set_error("elimrec_linear_x3_fwd: unsupported shape/alignment (K=%lld ldx=%lld)", (long long)K, (long long)ldx);
        return -2;
    }
    CUtensorMap tmA, tmBh, tmBl;
    if (make_map(&tmA, X, M, K, ldx, F_BM) != 0 || make_map(&tmBh, W_hi, 64, K, K, F_BN) != 0 ||
        make_map(&tmBl, W_lo, 64, K, K, F_BN) != 0) {
        elimrec_set_error("elimrec_linear_x3_fwd: cuTensorMapEncodeTiled failed");
        return -3;
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(linear_x3_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, X_SMEM);
        if (e != cudaSuccess) {
            elimrec_set_error("elimrec_linear_x3_fwd: shared-memory opt-in failed: %s", cudaGetErrorString(e));
            return -3;
        }
        configured = true;
    }
    const int num_kb = (int)((K + F_BK - 1) / F_BK);
    const unsigned grid = (unsigned)((M + F_BM - 1) / F_BM);
    linear_x3_fwd_kernel<<<grid, 192, X_SMEM, er_stream(stream)>>>(tmA, tmBh, tmBl, b, Y, ldy, (int)M, num_kb);
    ER_LAUNCH_CHECK();
    return 0;
}


ELIMREC_API int elimrec_fuse_heads_x3(int64_t rows, int n_heads, const float* O, int64_t ldo, const float* Wf_hi,
                                      const float* Wf_lo, const float* bf, const float* const* Ws_hi_host,
                                      const float* const* Ws_lo_host, const float* const* bs_host, float* F_out,
                                      float* const* S_out_host, elimrec_stream_t stream) {
    if (rows <= 0) return 0;
    ER_CHECK_ARG(n_heads >= 1 && n_heads <= ELIMREC_MAX_MODS, "n_heads out of range");
    const int64_t K = 64 * (1 + n_heads);
    if (ldo % 4 != 0 || ldo < K || !aligned16(O) || !aligned16(Wf_hi) || !aligned16(Wf_lo) || rows > 0x7fffffff) {
        elimrec_set_error("elimrec_fuse_heads_x3: unsupported shape/alignment");
        return -2;
    }
    CUtensorMap tmA, tmFh, tmFl;
    HeadMaps hm;
    HeadOut ho{};
    int bad = make_map(&tmA, O, rows, K, ldo, F_BM) | make_map(&tmFh, Wf_hi, 64, K, K, F_BN) | make_map(&tmFl, Wf_lo, 64, K, K, F_BN);
    ho.bias[0] = bf;
    ho.out[0] = F_out;
    for (int m = 0; m < ELIMREC_MAX_MODS; ++m) {
        const int mm = m < n_heads ? m : 0;   // unused maps still have to be valid descriptors
        bad |= make_map(&hm.hi[m], Ws_hi_host[mm], 64, 64, 64, F_BN) | make_map(&hm.lo[m], Ws_lo_host[mm], 64, 64, 64, F_BN);
        if (m < n_heads) {
            ho.bias[m + 1] = bs_host[m];
            ho.out[m + 1] = S_out_host[m];
        }
    }
    if (bad) {
        elimrec_set_error("elimrec_fuse_heads_x3: cuTensorMapEncodeTiled failed");
        return -3;
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fuse_heads_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM);
        if (e != cudaSuccess) {
            elimrec_set_error("elimrec_fuse_heads_x3: shared-memory opt-in failed: %s", cudaGetErrorString(e));
            return -3;
        }
        configured = true;
    }
    const unsigned grid = (unsigned)((rows + F_BM - 1) / F_BM);
    fuse_heads_x3_kernel<<<grid, 192, H_SMEM, er_stream(stream)>>>(tmA, tmFh, tmFl, hm, ho, (int)rows, n_heads);
    ER_LAUNCH_CHECK();
    return 0;
}


ELIMREC_API int elimrec_fuse_heads_x3_all(int64_t num_users, int64_t num_items, int n_heads, const float* O, int64_t ldo,
                                          const float* Wu_hi, const float* Wu_lo, const float* bu, const float* Wi_hi,
                                          const float* Wi_lo, const float* bi, const float* const* Ws_hi_host,
                                          const float* const* Ws_lo_host, const float* const* bs_host, float* F_out,
                                          float* const* S_out_host, elimrec_stream_t stream) {
    const int64_t N = num_users + num_items;
    if (N <= 0) return 0;
    ER_CHECK_ARG(n_heads >= 1 && n_heads <= ELIMREC_MAX_MODS, "n_heads out of range");
    const int64_t K = 64 * (1 + n_heads);
    if (ldo % 4 != 0 || ldo < K || !aligned16(O) || N > 0x7fffffff) {
        elimrec_set_error("elimrec_fuse_heads_x3_all: unsupported shape/alignment");
        return -2;
    }
    FuseAllMaps mp;
    FuseAllOut ho{};
    int bad = make_map(&mp.A, O, N, K, ldo, F_BM);
    bad |= make_map(&mp.Fh[0], Wu_hi, 64, K, K, F_BN) | make_map(&mp.Fl[0], Wu_lo, 64, K, K, F_BN);
    bad |= make_map(&mp.Fh[1], Wi_hi, 64, K, K, F_BN) | make_map(&mp.Fl[1], Wi_lo, 64, K, K, F_BN);
    ho.bias_f[0] = bu; ho.bias_f[1] = bi;
    ho.out[0] = F_out;
    for (int m = 0; m < ELIMREC_MAX_MODS; ++m) {
        const int mm = m < n_heads ? m : 0;
        bad |= make_map(&mp.Sh[m], Ws_hi_host[mm], 64, 64, 64, F_BN) | make_map(&mp.Sl[m], Ws_lo_host[mm], 64, 64, 64, F_BN);
        ho.bias_s[m] = bs_host[mm];
        if (m < n_heads) ho.out[m + 1] = S_out_host[m];
    }
    if (bad) {
        elimrec_set_error("elimrec_fuse_heads_x3_all: cuTensorMapEncodeTiled failed");
        return -3;
    }
    static bool configured = false;
    static int n_sm = 148;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fuse_heads_x3_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HA_SMEM);
        if (e != cudaSuccess) {
            elimrec_set_error("elimrec_fuse_heads_x3_all: shared-memory opt-in failed: %s", cudaGetErrorString(e));
            return -3;
        }
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    const int n_ut = (int)((num_users + F_BM - 1) / F_BM), n_it = (int)((num_items + F_BM - 1) / F_BM);
    const int n_tiles = n_ut + n_it;
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    fuse_heads_x3_all_kernel<<<grid, 320, HA_SMEM, er_stream(stream)>>>(mp, ho, (int)num_users, (int)N, n_heads, n_ut, n_tiles);
    ER_LAUNCH_CHECK();
    return 0;
}
