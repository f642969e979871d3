// pbn_fit.cu — COD scan of the Bittner predictor-set fitter on the device (see pbn_fit.cuh and include/pbn_b200.h).
//
// Replaces the O(G * C(G-1,3)) Python loop of gen/predictor_sets.py:41-78 (_gen_predictor_sets_gene: one pinv per
// candidate, one add_to_buff per candidate).  A block owns (target gene, first input gene a); its threads stride over
// the pairs (b, c) of later genes and walk every target row and every product of input rows, each keeping its own best
// `top_l` keys; the block merges them by `top_l` rounds of a block-wide minimum and a second small kernel merges the
// per-block lists of a gene.
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "pbn_b200.h"
#include "pbn_fit.cuh"

int pbn_fail_(int code, const std::string &msg);  // pbn_b200.cu
#define CKF(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            release();                                                                                   \
            return pbn_fail_(PBN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
        }                                                                                                \
    } while (0)

#define FIT_MAX_L 16
#define FIT_THREADS 256

struct FitView {
    int n_genes, n_samples, n_rows;
    const int *row_off;            // [G+1]
    const uint32_t *rows;          // [R]
    const uint16_t *cod_rank;      // [R][S+1]
    const unsigned long long *key_gt;  // [G] or NULL: keep only keys > key_gt[g]
    const unsigned long long *arr_lt;  // [G] or NULL: keep only candidates visited before arr_lt[g]
    const uint16_t *tie_rank_le;   // [G] or NULL: report rounding-tie candidates whose best case ranks <= this
    unsigned long long *top;       // [G][n_blk][top_l]
    unsigned long long *ties;      // [tie_cap][2] (key with the best-case rank, gene)
    unsigned long long *n_ties;
    long long tie_cap;
    int top_l;
};

__device__ __forceinline__ unsigned long long block_min(unsigned long long v, unsigned long long *sh) {
    for (int o = 16; o; o >>= 1) {
        const unsigned long long w = __shfl_xor_sync(0xFFFFFFFFu, v, o);
        v = w < v ? w : v;
    }
    __syncthreads();  // protects sh against the previous round's readers
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long m = sh[0];
    for (int i = 1; i < FIT_THREADS / 32; ++i) m = sh[i] < m ? sh[i] : m;
    return m;
}

// sorted insertion into a thread's top list (best[0] smallest); the caller has checked key < best[L-1]
__device__ __forceinline__ void top_insert(unsigned long long *best, int L, unsigned long long key) {
    int pos = L - 1;
    while (pos > 0 && best[pos - 1] > key) {
        best[pos] = best[pos - 1];
        --pos;
    }
    best[pos] = key;
}

// block merge: L rounds of "everyone offers its head, the minimum pops" (keys are unique)
__device__ __forceinline__ void top_merge(const unsigned long long *best, int L, unsigned long long *out, unsigned long long *s_red) {
    int head = 0;
    for (int r = 0; r < L; ++r) {
        const unsigned long long mine = head < L ? best[head] : FIT_KEY_NONE;
        const unsigned long long m = block_min(mine, s_red);
        if (m != FIT_KEY_NONE && mine == m) ++head;
        if (threadIdx.x == 0) out[r] = m;
    }
}

// grid (n_rem - 2, G): block (a, g) owns every triple a < b < c of the gene list with target g removed; its threads
// walk the (b, c) pairs with stride blockDim, so all lanes of a warp run the same trip count (+-1).
__global__ void __launch_bounds__(FIT_THREADS) k_fit_scan(FitView v) {
    extern __shared__ uint32_t smem[];
    uint32_t *s_rows = smem;                          // [R]
    int *s_off = (int *)(s_rows + v.n_rows);          // [G+1]
    __shared__ unsigned long long s_red[FIT_THREADS / 32];
    for (int i = threadIdx.x; i < v.n_rows; i += blockDim.x) s_rows[i] = v.rows[i];
    for (int i = threadIdx.x; i <= v.n_genes; i += blockDim.x) s_off[i] = v.row_off[i];
    __syncthreads();

    const int g = blockIdx.y, a = blockIdx.x, S = v.n_samples, L = v.top_l;
    const int n_rem = v.n_genes - 1;
    const unsigned long long key_gt = v.key_gt ? v.key_gt[g] : 0ull;
    const bool has_gt = v.key_gt != nullptr;
    const unsigned long long arr_lt = v.arr_lt ? v.arr_lt[g] : FIT_KEY_NONE;
    const bool want_ties = v.tie_rank_le != nullptr;
    const int tie_le = want_ties ? v.tie_rank_le[g] : -1;
    const int y0 = s_off[g], y1 = s_off[g + 1];
    const int ga = a + (a >= g);
    const int a0 = s_off[ga], a1 = s_off[ga + 1];

    unsigned long long best[FIT_MAX_L];
#pragma unroll
    for (int i = 0; i < FIT_MAX_L; ++i) best[i] = FIT_KEY_NONE;
    unsigned long long worst = FIT_KEY_NONE;  // best[L-1]

    // pairs (b, c), a < b < c < n_rem, in row-major order; this thread takes every blockDim-th one
    int b = a + 1, c = a + 2 + (int)threadIdx.x;
    for (;;) {
        while (c >= n_rem && b < n_rem - 1) {  // carry the overflow into the next rows
            const int over = c - n_rem;
            ++b;
            c = b + 1 + over;
        }
        if (b >= n_rem - 1) break;
        const int gb = b + (b >= g), gc = c + (c >= g);
        const int b0 = s_off[gb], b1 = s_off[gb + 1], c0 = s_off[gc], c1 = s_off[gc + 1];
        const int nb = b1 - b0, nc = c1 - c0;
        for (int y = y0; y < y1; ++y) {
            const uint32_t my = s_rows[y];
            const uint16_t *rank_row = v.cod_rank + (size_t)y * (S + 1);
            for (int ia = a0; ia < a1; ++ia)
                for (int ib = b0; ib < b1; ++ib)
                    for (int ic = c0; ic < c1; ++ic) {
                        const unsigned long long arr =
                            fit_arrival(a, b, c, y - y0, ((ia - a0) * nb + (ib - b0)) * nc + (ic - c0));
                        if (arr >= arr_lt) continue;
                        int k_lo, k_hi;
                        fit_eval(s_rows[ia], s_rows[ib], s_rows[ic], my, S, &k_lo, &k_hi);
                        k_lo = k_lo > S ? S : k_lo;  // beyond S errors the COD is negative and floored anyway
                        k_hi = k_hi > S ? S : k_hi;
                        const unsigned r_lo = __ldg(rank_row + k_lo);
                        if (k_lo != k_hi && __ldg(rank_row + k_hi) != r_lo) {
                            if (want_ties && (int)r_lo <= tie_le) {
                                const unsigned long long at = atomicAdd(v.n_ties, 1ull);
                                if ((long long)at < v.tie_cap) {
                                    v.ties[2 * at] = ((unsigned long long)r_lo << FIT_ARRIVAL_BITS) | arr;
                                    v.ties[2 * at + 1] = (unsigned long long)g;
                                }
                            }
                            continue;
                        }
                        const unsigned long long key = ((unsigned long long)r_lo << FIT_ARRIVAL_BITS) | arr;
                        if (key >= worst || (has_gt && key <= key_gt)) continue;
                        top_insert(best, L, key);
                        worst = best[L - 1];
                    }
        }
        c += blockDim.x;
    }
    top_merge(best, L, v.top + ((size_t)g * gridDim.x + blockIdx.x) * L, s_red);
}

// one block per gene: the per-block lists [n_blk][L] -> the gene's L best
__global__ void __launch_bounds__(FIT_THREADS) k_fit_merge(const unsigned long long *lists, int n_blk, int L, unsigned long long *out) {
    __shared__ unsigned long long s_red[FIT_THREADS / 32];
    const unsigned long long *mine = lists + (size_t)blockIdx.x * n_blk * L;
    unsigned long long best[FIT_MAX_L];
#pragma unroll
    for (int i = 0; i < FIT_MAX_L; ++i) best[i] = FIT_KEY_NONE;
    for (int i = threadIdx.x; i < n_blk * L; i += blockDim.x) {
        const unsigned long long key = mine[i];
        if (key < best[L - 1]) top_insert(best, L, key);
    }
    top_merge(best, L, out + (size_t)blockIdx.x * L, s_red);
}

// Test hook: the per-candidate arithmetic is __host__ __device__; this runs it on the host for n candidates so that the
// exact solver can be checked against rational arithmetic without a GPU.  Not used by the product.
extern "C" int pbn_fit_eval_host(const uint32_t *masks4, int64_t n, int32_t n_samples, int32_t *k_lo, int32_t *k_hi) {
    if (!masks4 || !k_lo || !k_hi || n_samples < 1 || n_samples > 32)
        return pbn_fail_(PBN_ERR_ARG, "pbn_fit_eval_host: bad arguments");
    for (int64_t i = 0; i < n; ++i) {
        int lo, hi;
        fit_eval(masks4[4 * i], masks4[4 * i + 1], masks4[4 * i + 2], masks4[4 * i + 3], n_samples, &lo, &hi);
        k_lo[i] = lo;
        k_hi[i] = hi;
    }
    return PBN_OK;
}

extern "C" int pbn_fit_scan_host(const PbnFitDesc *d, int32_t top_l, const uint64_t *key_gt, const uint64_t *arr_lt,
                                 const uint16_t *tie_rank_le, uint64_t *top_keys, uint64_t *tie_keys, int64_t tie_cap,
                                 int64_t *n_ties, float *kernel_ms) {
    if (!d || !top_keys) return pbn_fail_(PBN_ERR_ARG, "pbn_fit_scan_host: null argument");
    const int G = d->n_genes, S = d->n_samples;
    if (G < 1 || G > 4096) return pbn_fail_(PBN_ERR_ARG, "pbn_fit_scan_host: 1..4096 genes supported");
    if (S < 1 || S > 32) return pbn_fail_(PBN_ERR_ARG, "pbn_fit_scan_host: 1..32 samples supported (one 32-bit mask per row)");
    if (top_l < 1 || top_l > FIT_MAX_L) return pbn_fail_(PBN_ERR_ARG, "pbn_fit_scan_host: top_l must be in 1..16");
    if (!d->row_off || !d->rows || !d->cod_rank) return pbn_fail_(PBN_ERR_ARG, "pbn_fit_scan_host: null description array");
    if (tie_rank_le && (!tie_keys || !n_ties || tie_cap < 1))
        return pbn_fail_(PBN_ERR_ARG, "pbn_fit_scan_host: tie reporting needs a buffer");
    const int R = d->row_off[G];
    if (d->row_off[0] != 0 || R < G) return pbn_fail_(PBN_ERR_ARG, "pbn_fit_scan_host: row_off must start at 0 and give every gene a row");
    const uint32_t full = S >= 32 ? 0xFFFFFFFFu : ((1u << S) - 1u);
    for (int i = 0; i < G; ++i) {
        const int n = d->row_off[i + 1] - d->row_off[i];
        if (n < 1 || n > 16) return pbn_fail_(PBN_ERR_ARG, "pbn_fit_scan_host: 1..16 rows per gene supported");
    }
    for (int i = 0; i < R; ++i)
        if (d->rows[i] & ~full) return pbn_fail_(PBN_ERR_ARG, "pbn_fit_scan_host: row mask has bits beyond n_samples");
    for (int i = 0; i < R * (S + 1); ++i)
        if (d->cod_rank[i] > 1023) return pbn_fail_(PBN_ERR_ARG, "pbn_fit_scan_host: cod_rank must be < 1024");
    const size_t shmem = (size_t)R * 4 + (size_t)(G + 1) * 4;
    if (shmem > 200 * 1024) return pbn_fail_(PBN_ERR_ARG, "pbn_fit_scan_host: gene table exceeds shared memory");

    const int n_blk = G - 3;  // values of a: 0 .. n_rem-3
    if (n_blk < 1) {          // fewer than four genes: no triple of other genes exists
        for (int i = 0; i < G * top_l; ++i) top_keys[i] = FIT_KEY_NONE;
        if (n_ties) *n_ties = 0;
        if (kernel_ms) *kernel_ms = 0.f;
        return PBN_OK;
    }
    std::vector<void *> owned;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    auto release = [&]() {
        for (void *p : owned) cudaFree(p);
        owned.clear();
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        ev0 = ev1 = nullptr;
    };
    auto upload = [&](const void *src, size_t bytes, void **dst) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes ? bytes : 1);
        if (e != cudaSuccess) return e;
        owned.push_back(*dst);
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    FitView v{};
    v.n_genes = G;
    v.n_samples = S;
    v.n_rows = R;
    v.top_l = top_l;
    v.tie_cap = tie_rank_le ? tie_cap : 0;
    void *p = nullptr;
    CKF(upload(d->row_off, (size_t)(G + 1) * 4, &p));
    v.row_off = (const int *)p;
    CKF(upload(d->rows, (size_t)R * 4, &p));
    v.rows = (const uint32_t *)p;
    CKF(upload(d->cod_rank, (size_t)R * (S + 1) * 2, &p));
    v.cod_rank = (const uint16_t *)p;
    if (key_gt) {
        CKF(upload(key_gt, (size_t)G * 8, &p));
        v.key_gt = (const unsigned long long *)p;
    }
    if (arr_lt) {
        CKF(upload(arr_lt, (size_t)G * 8, &p));
        v.arr_lt = (const unsigned long long *)p;
    }
    if (tie_rank_le) {
        CKF(upload(tie_rank_le, (size_t)G * 2, &p));
        v.tie_rank_le = (const uint16_t *)p;
        CKF(cudaMalloc(&p, (size_t)tie_cap * 16));
        owned.push_back(p);
        v.ties = (unsigned long long *)p;
        CKF(cudaMalloc(&p, 8));
        owned.push_back(p);
        v.n_ties = (unsigned long long *)p;
        CKF(cudaMemset(p, 0, 8));
    }
    CKF(cudaMalloc(&p, (size_t)G * n_blk * top_l * 8));
    owned.push_back(p);
    v.top = (unsigned long long *)p;
    const size_t top_bytes = (size_t)G * top_l * 8;
    CKF(cudaMalloc(&p, top_bytes));
    owned.push_back(p);
    unsigned long long *d_final = (unsigned long long *)p;

    CKF(cudaFuncSetAttribute(k_fit_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    CKF(cudaEventCreate(&ev0));
    CKF(cudaEventCreate(&ev1));
    CKF(cudaEventRecord(ev0, 0));
    k_fit_scan<<<dim3(n_blk, G), FIT_THREADS, shmem>>>(v);
    CKF(cudaGetLastError());
    k_fit_merge<<<G, FIT_THREADS>>>(v.top, n_blk, top_l, d_final);
    CKF(cudaGetLastError());
    CKF(cudaEventRecord(ev1, 0));
    CKF(cudaMemcpy(top_keys, d_final, top_bytes, cudaMemcpyDeviceToHost));
    if (kernel_ms) CKF(cudaEventElapsedTime(kernel_ms, ev0, ev1));
    if (tie_rank_le) {
        unsigned long long n = 0;
        CKF(cudaMemcpy(&n, v.n_ties, 8, cudaMemcpyDeviceToHost));
        *n_ties = (int64_t)n;
        const unsigned long long got = n < (unsigned long long)tie_cap ? n : (unsigned long long)tie_cap;
        if (got) CKF(cudaMemcpy(tie_keys, v.ties, (size_t)got * 16, cudaMemcpyDeviceToHost));
    }
    release();
    return PBN_OK;
}
