// GPU BOX: what one tcgen05.mma (cta_group::1, M = 128, both operands from shared memory) costs as a function of N, for
// kind::f16 (K = 16) and kind::tf32 (K = 8), K-major 128-byte-swizzle operands and (tf32) MN-major operands - the shapes
// the ranking / 3xTF32 kernels issue.  One thread issues `reps` MMAs back to back on the same operand tiles, commits, waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe tools/probes/mma_probe.cu && ./mma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nW_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra W_DONE;\nbra W_LOOP;\nW_DONE:\n}\n"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
template <int KIND>   // 0 = f16, 1 = tf32
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: f16 K-major sw128;  1: tf32 K-major sw128;  2: tf32 MN-major (sw128, 32-byte atoms)
template <int MODE>
__global__ void __launch_bounds__(128, 1) probe(int N, int reps, int n_acc, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t d = tbase;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 32 * 1024;
        uint32_t idesc;
        if (MODE == 0) idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        else idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((MODE == 2 ? 1u : 0u) << 15) | ((MODE == 2 ? 1u : 0u) << 16) |
                     ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        uint64_t da[4], db[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (MODE == 2) { da[k] = desc(a0 + k * 1024, 4096, 512, 1); db[k] = desc(b0 + k * 1024, 4096, 512, 1); }
            else { da[k] = desc(a0 + k * 32, 16, 1024, 2); db[k] = desc(b0 + k * 32, 16, 1024, 2); }
        }
        const uint32_t amask = (uint32_t)n_acc - 1;      // n_acc is a power of two
        const long long t0 = clock64();
        for (int r = 0; r < reps; r += 4) {              // descriptors precomputed, four MMAs per trip: nothing but the issue
            const uint32_t dd = d + ((uint32_t)(r >> 2) & amask) * (uint32_t)N;
            mma<MODE == 0 ? 0 : 1>(dd, da[0], db[0], idesc, r >= 4 * n_acc);
            mma<MODE == 0 ? 0 : 1>(dd, da[1], db[1], idesc, 1);
            mma<MODE == 0 ? 0 : 1>(dd, da[2], db[2], idesc, 1);
            mma<MODE == 0 ? 0 : 1>(dd, da[3], db[3], idesc, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(d), "r"(512) : "memory");
}

template <int MODE>
void run(const char* name, int blocks) {
    long long* out;
    cudaMalloc(&out, 1024 * sizeof(long long));
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 4096;
    for (int N : {64, 128, 256}) {
        for (int n_acc : {1, 512 / N}) {
            probe<MODE><<<blocks, 128, 100 * 1024>>>(N, reps, n_acc, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s N=%d: %s\n", name, N, cudaGetErrorString(e)); return; }
            long long h[1024];
            cudaMemcpy(h, out, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
            double s = 0;
            for (int i = 0; i < blocks; ++i) s += (double)h[i];
            printf("%-28s N=%3d accumulators=%d CTAs=%3d: %.1f cycles per MMA (math floor %d)\n", name, N, n_acc, blocks,
                   s / blocks / reps, 128 * N / 256);
        }
    }
    cudaFree(out);
}

int main() {
    for (int blocks : {1, 148}) {
        run<0>("f16  K-major  (K=16)", blocks);
        run<1>("tf32 K-major  (K=8)", blocks);
        run<2>("tf32 MN-major (K=8)", blocks);
    }
    return 0;
}
