// Column-sharded ("feature-sharded") multi-GPU step - the glue kernels around the exchanges (elimrec_b200/colshard.py).
//
// The linear schedule propagates ONE 64-wide slab, and propagation acts on every column independently: rank r of G owns
// columns [r w, (r+1) w), w = 64 / G, of both embedding tables and of every propagation / gradient slab, the CSR is
// replicated, and a propagation layer needs NO communication at all.  What crosses NVLink per step is only what the loss
// reads and writes at the 3B instance rows of each rank's own batch:
//   forward   rank r computes its w columns of (mean_k p_k, parity mean) at the instance rows of EVERY rank's batch (pack),
//             one all-to-all hands each rank all 64 columns of ITS 3B rows (unpack); modality GEMMs, fusion, heads and the
//             BPR losses then run on the rank's own batch exactly as on one GPU;
//   backward  the two seed vectors of the rank's 3B rows are cut into G column slices (seed_pack), one all-to-all gives every
//             rank its columns of all G 3B rows (seed_scatter), then the backward chain and Adam run on the rank's columns.
// All four kernels are element-wise; a thread owns one (row, column).
#include "common.cuh"

namespace {

struct CsLayers {
    int n;
    const float* user[ELIMREC_MAX_LAYERS + 1];
    const float* item[ELIMREC_MAX_LAYERS + 1];
    long long user_ld[ELIMREC_MAX_LAYERS + 1];
    long long item_ld[ELIMREC_MAX_LAYERS + 1];
};

// out[j, 0:w] = scale * sum_k p_k[node_j, 0:w];  out[j, w:2w] = scale * sum over the parity layers (users: even k, items: odd k)
__global__ void cs_pack_kernel(long long n_rows, const int* __restrict__ rows, int num_users, CsLayers lay, float scale, int w,
                               float* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long j = t / w;
    const int c = (int)(t - j * w);
    if (j >= n_rows) return;
    const long long node = (rows != nullptr) ? (long long)__ldg(rows + j) : j;
    const bool is_user = node < num_users;
    const long long r = is_user ? node : node - num_users;
    const int par = is_user ? 0 : 1;
    float sa = 0.f, sp = 0.f;
    for (int k = 0; k < lay.n; ++k) {
        const float v = is_user ? __ldg(lay.user[k] + r * lay.user_ld[k] + c) : __ldg(lay.item[k] + r * lay.item_ld[k] + c);
        sa = (k == 0) ? v : sa + v;
        if ((k & 1) == par) sp += v;
    }
    out[j * 2 * w + c] = sa * scale;
    out[j * 2 * w + w + c] = sp * scale;
}

// recv [G][n][2w] -> O[j, q w + c] = recv[q][j][c];  O[j, 64 (1+m) + q w + c] += recv[q][j][w + c]  for m < n_mod
__global__ void cs_unpack_kernel(int G, long long n, int w, const float* __restrict__ recv, int n_mod, float* __restrict__ O,
                                 long long ldo) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long j = t >> 6;
    const int col = (int)(t & 63);
    if (j >= n) return;
    const int q = col / w, c = col - q * w;
    const float* src = recv + ((long long)q * n + j) * 2 * w;
    const float all = src[c], par = src[w + c];
    float* o = O + j * ldo;
    o[col] = all;
    for (int m = 0; m < n_mod; ++m) o[64 * (m + 1) + col] += par;
}

// send[q][j][c] = scale * sum_b dO[j, 64 b + q w + c];  send[q][j][w + c] = scale * dO[j, q w + c]
__global__ void cs_seed_pack_kernel(long long n, int G, int w, const float* __restrict__ dO, long long ldo, int n_mod, float scale,
                                    float* __restrict__ send) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long j = t >> 6;
    const int col = (int)(t & 63);
    if (j >= n) return;
    const float* s = dO + j * ldo;
    const float b = s[col];
    float a = b;
    for (int m = 0; m < n_mod; ++m) a += s[64 * (m + 1) + col];
    const int q = col / w, c = col - q * w;
    float* d = send + ((long long)q * n + j) * 2 * w;
    d[c] = scale * a;
    d[w + c] = scale * b;
}

// GA[node_j, c] += recv[j][c];  GB[node_j, c] += recv[j][w + c]   (atomic: nodes repeat inside and across batches)
__global__ void cs_seed_scatter_kernel(long long n_all, const int* __restrict__ rows, int w, const float* __restrict__ recv,
                                       float* __restrict__ GA, float* __restrict__ GB, long long ldg) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long j = t / w;
    const int c = (int)(t - j * w);
    if (j >= n_all) return;
    const long long node = __ldg(rows + j);
    atomicAdd(GA + node * ldg + c, recv[j * 2 * w + c]);
    atomicAdd(GB + node * ldg + c, recv[j * 2 * w + w + c]);
}

// triples of every rank's batch, T [G][3][B] (users | pos | neg per rank) -> node ids rows[G][3][B] (items offset by U), and the
// instance-row masks over ALL batches
__global__ void cs_inst_rows_kernel(long long n, int B, const long long* __restrict__ T, int num_users, int* __restrict__ rows,
                                    unsigned char* __restrict__ mask, unsigned char* __restrict__ mask2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = (int)((i / B) % 3);
    const int node = (int)__ldg(T + i) + (k == 0 ? 0 : num_users);
    rows[i] = node;
    if (mask != nullptr) mask[node] = 1;
    if (mask2 != nullptr) mask2[node] = 1;
}

inline unsigned cs_blocks(long long threads) { return (unsigned)((threads + 255) / 256); }

}  // namespace

ELIMREC_API int elimrec_cs_pack(int64_t n_rows, const int32_t* rows, int32_t num_users, const elimrec_lin_layers_t* layers,
                                float scale, int w, float* out, elimrec_stream_t stream) {
    ER_CHECK_ARG(layers != nullptr && layers->n >= 1 && layers->n <= ELIMREC_MAX_LAYERS + 1, "1..MAX_LAYERS+1 layer tables");
    ER_CHECK_ARG(w >= 1 && w <= 64 && 64 % w == 0 && out != nullptr, "w must divide 64");
    if (n_rows <= 0) return 0;
    CsLayers lay{};
    lay.n = layers->n;
    for (int k = 0; k < lay.n; ++k) {
        lay.user[k] = layers->user[k]; lay.item[k] = layers->item[k];
        lay.user_ld[k] = layers->user_ld[k]; lay.item_ld[k] = layers->item_ld[k];
    }
    cs_pack_kernel<<<cs_blocks(n_rows * w), 256, 0, er_stream(stream)>>>(n_rows, rows, num_users, lay, scale, w, out);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_cs_unpack(int world, int64_t n, int w, const float* recv, int n_mod, float* O, int64_t ldo,
                                  elimrec_stream_t stream) {
    ER_CHECK_ARG(world >= 1 && w * world == 64 && recv != nullptr && O != nullptr, "w * world must be 64");
    ER_CHECK_ARG(n_mod >= 0 && n_mod <= ELIMREC_MAX_MODS && ldo >= 64 * (1 + n_mod), "bad output shape");
    if (n <= 0) return 0;
    cs_unpack_kernel<<<cs_blocks(n * 64), 256, 0, er_stream(stream)>>>(world, n, w, recv, n_mod, O, ldo);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_cs_seed_pack(int64_t n, int world, int w, const float* dO, int64_t ldo, int n_mod, float scale, float* send,
                                     elimrec_stream_t stream) {
    ER_CHECK_ARG(world >= 1 && w * world == 64 && dO != nullptr && send != nullptr, "w * world must be 64");
    ER_CHECK_ARG(n_mod >= 0 && n_mod <= ELIMREC_MAX_MODS, "n_mod out of range");
    if (n <= 0) return 0;
    cs_seed_pack_kernel<<<cs_blocks(n * 64), 256, 0, er_stream(stream)>>>(n, world, w, dO, ldo, n_mod, scale, send);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_cs_seed_scatter(int64_t n_all, const int32_t* rows, int w, const float* recv, float* GA, float* GB,
                                        int64_t ldg, elimrec_stream_t stream) {
    ER_CHECK_ARG(rows != nullptr && recv != nullptr && GA != nullptr && GB != nullptr && w >= 1 && w <= 64, "bad arguments");
    if (n_all <= 0) return 0;
    cs_seed_scatter_kernel<<<cs_blocks(n_all * w), 256, 0, er_stream(stream)>>>(n_all, rows, w, recv, GA, GB, ldg);
    ER_LAUNCH_CHECK();
    return 0;
}

ELIMREC_API int elimrec_cs_inst_rows(int world, int B, const int64_t* triples, int32_t num_users, int32_t* rows, int64_t n_nodes,
                                     uint8_t* mask, uint8_t* mask2, elimrec_stream_t stream) {
    ER_CHECK_ARG(world >= 1 && B >= 0 && triples != nullptr && rows != nullptr, "bad arguments");
    cudaStream_t st = er_stream(stream);
    for (uint8_t* m : {mask, mask2}) {
        if (m != nullptr && n_nodes > 0 && cudaMemsetAsync(m, 0, (size_t)n_nodes, st) != cudaSuccess) {
            elimrec_set_error("elimrec_cs_inst_rows: memset failed");
            return -3;
        }
    }
    const long long n = (long long)world * 3 * B;
    if (n == 0) return 0;
    cs_inst_rows_kernel<<<cs_blocks(n), 256, 0, st>>>(n, B, (const long long*)triples, num_users, rows, mask, mask2);
    ER_LAUNCH_CHECK();
    return 0;
}
