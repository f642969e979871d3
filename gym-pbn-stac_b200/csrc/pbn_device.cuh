// pbn_device.cuh — device-side building blocks of the sm_100a PB(C)N kernels.
//
//   * Philox4x32-10 counter-based stream, one sequential stream per env (key = seed,
//     counter = (block, epoch, env_lo, env_hi)); 31-bit integer thresholds for every decision.
//   * replay stream: recorded draws of the reference's RNGs, float64 compares exactly as
//     common/node.py:37-38 and bittner/base.py:94-97 do them.
//   * bit-packed state in shared memory, one column per thread (word w of thread t at
//     s[w*blockDim + t]): dynamic bit indexing is a conflict-free LDS, never a register select.
//   * packed network image staged once per block into shared memory.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pbn_b200.h"

typedef uint32_t u32;
typedef uint64_t u64;

// ----------------------------------------------------------------------------------------------- views
// Packed network image (what pbn_net_create builds).  All offsets are in bytes from `blob`.
//   PRED: thr  u32 [N][4*tsq_stride]  cumulative 31-bit thresholds (ts used), padded with 0x80000000 (never <= r31)
//         rec  uint2 [N][fmax]  .x = four u8 node indices (in0 | in1<<8 | in2<<16 | self<<24), .y = 16-bit LUT
//   TT  : node uint2 [N]        .x = table offset, .y = in_off | k<<16
//         in   u16 []           input node indices (first = MSB of the table index)
//         thr  u32 []           31-bit thresholds ceil(P * 2^31)
struct NetView {
    int kind, n, first, w32;
    int ts, fmax;
    int tsq_stride;  // quads per threshold row: odd, so that divergent LDS.128 reads spread over all 8 bank groups
    int blob_bytes;       // staged by every kernel: the image without the fast-path records
    int blob_fast_bytes;  // image including the 16-byte predictor records (fast asynchronous paths), == blob_bytes when absent
    int off_thr, off_rec, off_node, off_in;
    int off_rec16;        // 0: no 16-byte records
    const unsigned char *blob;  // device
    const u32 *thr_dev;         // TT: non-null when the threshold table is read from global memory instead of the staged image
    // float64 side tables for replay mode (device, reference form)
    const int *pr_off;
    const double *pr_cum, *pr_codsum;
    const double *tt_prob;
    // bit-sliced synchronous mode (predictor nets, fmax <= 5): u32 [N*fmax][16], entry k = all-ones iff LUT bit k is set
    const u32 *lutmask;
};

// Packed env image: cubes as (care, value) word pairs.
struct EnvView {
    int kind, horizon, max_inner, force, dedup, control_write, n_control;
    int successful_reward, wrong_attractor_cost;
    int n_att, n_cubes, tgt_first, n_tgt;
    int img_bytes;               // bytes of [att_off (n_att+1 ints, padded to 8)] [cubes: n_cubes*w32*2 u32]
    int off_cubes;
    const unsigned char *img;    // device
    const double *gamma_pow;     // device, self-triggering envs: gamma**i for i < n_gamma
    int n_gamma, max_interval;
};

struct DrawView {
    int mode;
    u32 epoch;
    u32 seed_lo, seed_hi;
    const int *ints;
    const double *dbls;
    long long int_stride, dbl_stride;
    long long *used;
    const u32 *epoch_ptr;  // optional: the launch's epoch is epoch + *epoch_ptr (CUDA-graph replays, see PbnDraws.epoch_dev)
    // Philox round keys (rk[2r] = seed_lo + r*0x9E3779B9, rk[2r+1] = seed_hi + r*0xBB67AE85), computed once on the host:
    // read from the kernel-parameter bank they are immediate operands of the round's XOR instead of ten key updates per block
    u32 rk[20];
};

// ----------------------------------------------------------------------------------------------- Philox
__device__ __forceinline__ void philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32 &o0, u32 &o1,
                                              u32 &o2, u32 &o3) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        u64 p0 = (u64)0xD2511F53u * c0;
        u64 p1 = (u64)0xCD9E8D57u * c2;
        u32 n0 = (u32)(p1 >> 32) ^ c1 ^ k0;
        u32 n2 = (u32)(p0 >> 32) ^ c3 ^ k1;
        c1 = (u32)p1;
        c3 = (u32)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

#define PBN_PERTURB_BLOCK0 0x80000000u

// Philox4x32-10 with the round keys taken from the launch's DrawView (a kernel parameter)
__device__ __forceinline__ void philox4x32_10_rk(u32 c0, u32 c1, u32 c2, u32 c3, const DrawView &dv, u32 &o0, u32 &o1, u32 &o2,
                                                 u32 &o3) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        u64 p0 = (u64)0xD2511F53u * c0;
        u64 p1 = (u64)0xCD9E8D57u * c2;
        u32 n0 = (u32)(p1 >> 32) ^ c1 ^ dv.rk[2 * r];
        u32 n2 = (u32)(p0 >> 32) ^ c3 ^ dv.rk[2 * r + 1];
        c1 = (u32)p1;
        c3 = (u32)p0;
        c0 = n0;
        c2 = n2;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

template <int MODE>
struct Draw;

template <>
struct Draw<PBN_DRAW_PHILOX> {
    u32 k0, k1, blk, c1, c2, c3;
    u32 b0, b1, b2, b3;
    int have;
    u32 base;  // first block index of this stream: 0 = the env's update stream, PBN_PERTURB_BLOCK0 = its perturbation stream
    __device__ __forceinline__ void init(const DrawView &dv, long long /*local*/, long long env_id) {
        k0 = dv.seed_lo; k1 = dv.seed_hi;
        blk = base = 0; c1 = dv.epoch; c2 = (u32)env_id; c3 = (u32)((u64)env_id >> 32);
        have = 0;
        b0 = b1 = b2 = b3 = 0;
    }
    // second stream of the same env (SSD perturbation gaps): same key and counter words, block indices from 2^31 on, so
    // the update stream is consumed at a fixed rate (two draws per update) whatever the perturbations need
    __device__ __forceinline__ void init_perturb(const DrawView &dv, long long local, long long env_id) {
        init(dv, local, env_id);
        blk = base = PBN_PERTURB_BLOCK0;
    }
    __device__ __forceinline__ u32 next() {
        if (have == 0) {
            philox4x32_10(blk, c1, c2, c3, k0, k1, b0, b1, b2, b3);
            blk++;
            have = 4;
        }
        u32 r = b0;
        b0 = b1; b1 = b2; b2 = b3;
        have--;
        return r;
    }
    __device__ __forceinline__ u32 next_rk(const DrawView &dv) {  // next() with the host-computed round keys
        if (have == 0) {
            philox4x32_10_rk(blk, c1, c2, c3, dv, b0, b1, b2, b3);
            blk++;
            have = 4;
        }
        u32 r = b0;
        b0 = b1; b1 = b2; b2 = b3;
        have--;
        return r;
    }
    __device__ __forceinline__ u32 consumed() const { return (blk - base) * 4u - (u32)have; }  // draws taken so far
    // uniform integer in [lo, lo+n): random.randint(lo, lo+n-1)
    __device__ __forceinline__ int randint(int lo, int n) { return lo + (int)__umulhi(next(), (u32)n); }
    __device__ __forceinline__ void done(const DrawView &dv, long long local) {
        if (dv.used) { dv.used[2 * local] = consumed(); dv.used[2 * local + 1] = 0; }
    }
    // position of the stream = number of draws consumed; seek() re-creates the state at a position, so an env can be
    // handed from one thread to another (block-level compaction of divergent inner loops)
    __device__ __forceinline__ void tell(u32 &a, u32 &b) const { a = consumed(); b = 0; }
    __device__ __forceinline__ void seek(const DrawView &dv, long long local, long long env_id, u32 a, u32 /*b*/) {
        init(dv, local, env_id);
        blk = a >> 2;
        const u32 r = a & 3u;
        if (r) {
            u32 x0, x1, x2, x3;
            philox4x32_10(blk, c1, c2, c3, k0, k1, x0, x1, x2, x3);
            blk++;
            have = 4 - (int)r;
            b0 = r == 1 ? x1 : (r == 2 ? x2 : x3);
            b1 = r == 1 ? x2 : x3;
            b2 = x3;
        }
    }
};

template <>
struct Draw<PBN_DRAW_REPLAY> {
    const int *ip;
    const double *dp;
    long long ni, nd;
    __device__ __forceinline__ void init(const DrawView &dv, long long local, long long /*env_id*/) {
        ip = dv.ints + local * dv.int_stride;
        dp = dv.dbls + local * dv.dbl_stride;
        ni = nd = 0;
    }
    __device__ __forceinline__ void init_perturb(const DrawView &dv, long long local, long long env_id) { init(dv, local, env_id); }
    __device__ __forceinline__ u32 next() { return 0u; }  // never drawn from in replay mode (flips replay float64 draws)
    __device__ __forceinline__ int randint(int /*lo*/, int /*n*/) { ni++; return *ip++; }  // recorded result
    __device__ __forceinline__ double dbl() { nd++; return *dp++; }
    __device__ __forceinline__ void done(const DrawView &dv, long long local) {
        if (dv.used) { dv.used[2 * local] = ni; dv.used[2 * local + 1] = nd; }
    }
    __device__ __forceinline__ void tell(u32 &a, u32 &b) const { a = (u32)ni; b = (u32)nd; }
    __device__ __forceinline__ void seek(const DrawView &dv, long long local, long long env_id, u32 a, u32 b) {
        init(dv, local, env_id);
        ip += a; dp += b; ni = a; nd = b;
    }
};

// ----------------------------------------------------------------------------------------------- geometric gap
// Number of failures before the next success of a Bernoulli(p) process, from one 32-bit draw:
//   u = ((r>>9)+0.5)/2^23,  G = trunc(log2(u) * inv),  inv = 1/log2(1-p).
// log2 is a fixed degree-7 polynomial evaluated with IEEE single fma only (no MUFU), so the CPU oracle
// (oracle/pbn_oracle.c: orc_geom) reproduces it bit for bit.
__device__ __forceinline__ float log2f_poly(float x) {
    u32 b = __float_as_uint(x);
    int e = (int)(b >> 23) - 127;
    float m = __uint_as_float((b & 0x007FFFFFu) | 0x3F800000u);
    if (m > 1.41421356f) { m = __fmul_rn(m, 0.5f); e += 1; }
    float t = __fsub_rn(m, 1.0f);
    float p = -1.427597404e-01f;
    p = __fmaf_rn(p, t, 2.326525748e-01f);
    p = __fmaf_rn(p, t, -2.492718250e-01f);
    p = __fmaf_rn(p, t, 2.872888744e-01f);
    p = __fmaf_rn(p, t, -3.602251709e-01f);
    p = __fmaf_rn(p, t, 4.809167087e-01f);
    p = __fmaf_rn(p, t, -7.213529348e-01f);
    p = __fmaf_rn(p, t, 1.442695022e+00f);
    return __fmaf_rn(p, t, (float)e);
}
// u = (k + 0.5) * 2^-23 for the top 23 bits k of the draw; one fma: every intermediate is exact, so this is the oracle's
// (k + 0.5f) * 2^-23 bit for bit
__device__ __forceinline__ float gap_uniform(u32 r) { return __fmaf_rn((float)(r >> 9), 1.0f / 8388608.0f, 1.0f / 16777216.0f); }
__device__ __forceinline__ u32 geom_gap(u32 r, float inv) {
    float u = gap_uniform(r);
    float g = __fmul_rn(log2f_poly(u), inv);
    if (!(g < 33554432.0f)) return 33554432u;  // gaps are capped at 2^25 so a warp prefix sum over a round's 64 gaps fits 32 bits
    return (u32)g;                              // g >= 0: truncation
}

// Shortcut for the same gap: hardware lg2.approx instead of the polynomial.  It is taken only when the result is provably the
// polynomial's: the two logarithms differ by less than a margin dlt after scaling, so whenever the approximate g lies at least dlt
// away from an integer both truncate to the same gap; otherwise *ok is false and the caller evaluates geom_gap.  The host
// does not rely on the documented error bound of lg2.approx alone: before a value of `inv` is used with the shortcut,
// k_geom_verify compares the two functions on all 2^23 possible inputs (pbn_b200.cu: geom_shortcut_delta).
__device__ __forceinline__ u32 geom_gap_approx(u32 r, float inv, float half_minus_dlt, bool *ok) {
    const float u = gap_uniform(r);
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u));
    const float g = __fmul_rn(lg, inv);
    const float fr = __fsub_rn(g, truncf(g));  // exact; 0 for g >= 2^23 (then never ok)
    *ok = fabsf(__fsub_rn(fr, 0.5f)) < half_minus_dlt;  // the caller passes 0.5 - dlt
    return (u32)g;
}

// ----------------------------------------------------------------------------------------------- state column
#define PBN_BLOCK 256  // every kernel runs 256-thread blocks, so a column's word stride is a compile-time constant
struct Col {
    u32 *s;  // &state_smem[threadIdx.x]; word w of this thread lives at s[w * PBN_BLOCK]
    static constexpr int stride = PBN_BLOCK;
    __device__ __forceinline__ u32 *wp(u32 pos) const {  // address of the word holding bit `pos`: (pos & ~31) * (4*256/32) bytes on
        return reinterpret_cast<u32 *>(reinterpret_cast<char *>(s) + ((pos & ~31u) << 5));
    }
    __device__ __forceinline__ u32 word(int w) const { return s[w * stride]; }
    __device__ __forceinline__ void set_word(int w, u32 v) const { s[w * stride] = v; }
    __device__ __forceinline__ u32 bit(u32 pos) const { return (*wp(pos) >> (pos & 31u)) & 1u; }
    __device__ __forceinline__ void flip(u32 pos) const { *wp(pos) ^= 1u << (pos & 31u); }
    __device__ __forceinline__ void put(u32 pos, u32 v) const {
        u32 *p = wp(pos);
        u32 m = 1u << (pos & 31u);
        *p = (*p & ~m) | (v ? m : 0u);
    }
};

// ----------------------------------------------------------------------------------------------- node updates
// bittner/base.py:89-119 Node.Predstep.  `blob` is the shared-memory copy of the network image.
// TQ = number of threshold quads per node known at compile time (1: up to 5 predictors), 0 = read nv.ts at run time
template <int MODE, int TQ>
__device__ __forceinline__ u32 pred_next(const NetView &nv, const unsigned char *blob, const Col &st, int i, Draw<MODE> &d,
                                         u32 word = 0, bool have_word = false) {
    int j;
    if constexpr (MODE == PBN_DRAW_PHILOX) {
        u32 r = (have_word ? word : d.next()) >> 1;
        const int nq = TQ > 0 ? TQ : (nv.ts >> 2);
        const uint4 *thr = reinterpret_cast<const uint4 *>(blob + nv.off_thr) + i * nv.tsq_stride;
        if (nq == 1) {
            const uint4 t = thr[0];
            j = (t.x <= r) + (t.y <= r) + (t.z <= r) + (t.w <= r);
        } else if (nq <= 4) {
            // rows of two to four quads start with a quad holding the last threshold of each quad (padding = never <= r):
            // find the quad, then the position inside it — two loads however many predictors the node has
            const uint4 m = thr[0];
            int q = (m.x <= r) + (m.y <= r) + (m.z <= r) + (m.w <= r);
            q = q < nq - 1 ? q : nq - 1;
            const uint4 t = thr[1 + q];
            j = 4 * q + (t.x <= r) + (t.y <= r) + (t.z <= r) + (t.w <= r);
        } else {  // more than 17 predictors per node: flat row, linear scan
            j = 0;
            for (int q = 0; q < nq; q++) {
                const uint4 t = thr[q];
                j += (t.x <= r) + (t.y <= r) + (t.z <= r) + (t.w <= r);
            }
        }
    } else {
        double r = d.dbl() * nv.pr_codsum[i];  // base.py:94
        int q0 = nv.pr_off[i], q1 = nv.pr_off[i + 1];
        j = q1 - q0 - 1;  // falls through to the last predictor, base.py:95-97
        for (int k = q0; k < q1; k++)
            if (nv.pr_cum[k] > r) { j = k - q0; break; }
    }
    const uint2 rec = reinterpret_cast<const uint2 *>(blob + nv.off_rec)[i * nv.fmax + j];
    // four gathers: byte k of rec.x is a node index; its word sits (idx & 0xE0) << 5 bytes into the column and the
    // funnel-style shift below only looks at the low 5 bits of its count, so no per-field masking is needed
    const u32 p0 = rec.x, p1 = rec.x >> 8, p2 = rec.x >> 16, p3 = rec.x >> 24;
    const char *base = reinterpret_cast<const char *>(st.s);
    const u32 w0 = *reinterpret_cast<const u32 *>(base + ((p0 & 0xE0u) << 5));
    const u32 w1 = *reinterpret_cast<const u32 *>(base + ((p1 & 0xE0u) << 5));
    const u32 w2 = *reinterpret_cast<const u32 *>(base + ((p2 & 0xE0u) << 5));
    const u32 w3 = *reinterpret_cast<const u32 *>(base + ((p3 & 0xE0u) << 5));
    u32 idx = __funnelshift_r(w0, 0, p0) & 1u;
    idx = idx * 2u + (__funnelshift_r(w1, 0, p1) & 1u);
    idx = idx * 2u + (__funnelshift_r(w2, 0, p2) & 1u);
    idx = idx * 2u + (__funnelshift_r(w3, 0, p3) & 1u);
    return (rec.y >> idx) & 1u;
}

// common/node.py:31-38 Node.compute_next_value
template <int MODE>
__device__ __forceinline__ u32 tt_next(const NetView &nv, const unsigned char *blob, const Col &st, int i, Draw<MODE> &d,
                                       u32 word = 0, bool have_word = false) {
    uint2 nr = reinterpret_cast<const uint2 *>(blob + nv.off_node)[i];
    const unsigned short *in = reinterpret_cast<const unsigned short *>(blob + nv.off_in) + (nr.y & 0xFFFF);
    int k = (int)(nr.y >> 16);
    u32 idx = 0;
    for (int q = 0; q < k; q++) idx = (idx << 1) | st.bit(in[q]);
    if constexpr (MODE == PBN_DRAW_PHILOX) {
        const u32 thr = nv.thr_dev ? __ldg(nv.thr_dev + nr.x + idx) : reinterpret_cast<const u32 *>(blob + nv.off_thr)[nr.x + idx];
        return (((have_word ? word : d.next()) >> 1) < thr) ? 1u : 0u;
    } else {
        return (d.dbl() < nv.tt_prob[nr.x + idx]) ? 1u : 0u;  // u < p, node.py:37-38
    }
}

template <int NET, int MODE, int TQ>
__device__ __forceinline__ u32 node_next(const NetView &nv, const unsigned char *blob, const Col &st, int i, Draw<MODE> &d,
                                         u32 word = 0, bool have_word = false) {
    if constexpr (NET == PBN_NET_PRED) return pred_next<MODE, TQ>(nv, blob, st, i, d, word, have_word);
    else return tt_next<MODE>(nv, blob, st, i, d, word, have_word);
}

// the same update from two given 32-bit words of the env's update stream (Philox mode): wa picks the node, wb decides
template <int NET, int TQ = 0>
__device__ __forceinline__ void micro_step_words(const NetView &nv, const unsigned char *blob, const Col &st, u32 wa, u32 wb,
                                                 Draw<PBN_DRAW_PHILOX> &d) {
    const int i = nv.first + (int)__umulhi(wa, (u32)(nv.n - nv.first));
    const u32 v = node_next<NET, PBN_DRAW_PHILOX, TQ>(nv, blob, st, i, d, wb, true);
    st.put(i, v);
}

// one asynchronous update: PBN.step common/pbn.py:88-92 / PBCN.step common/pbcn.py:59-61 / Graph.step base.py:306-312
template <int NET, int MODE, int TQ = 0>
__device__ __forceinline__ void micro_step(const NetView &nv, const unsigned char *blob, const Col &st, Draw<MODE> &d) {
    int i = d.randint(nv.first, nv.n - nv.first);
    u32 v = node_next<NET, MODE, TQ>(nv, blob, st, i, d);
    st.put(i, v);
}

// Graph.synch_step (perturbations off) base.py:300-303: every node from the OLD state, in node order
template <int NET, int MODE, int TQ = 0>
__device__ __forceinline__ void sync_step(const NetView &nv, const unsigned char *blob, const Col &st, const Col &tmp,
                                          Draw<MODE> &d) {
    u32 acc = 0;
    for (int i = 0; i < nv.n; i++) {
        acc |= node_next<NET, MODE, TQ>(nv, blob, st, i, d) << (i & 31);
        if ((i & 31) == 31 || i == nv.n - 1) { tmp.set_word(i >> 5, acc); acc = 0; }
    }
    for (int w = 0; w < nv.w32; w++) st.set_word(w, tmp.word(w));
}

// The same step when the node count is a multiple of four (Philox mode): node i of step t takes word t*n + i of the
// env's stream, so every Philox block serves four consecutive nodes and no buffer bookkeeping is needed.
template <int NET, int TQ = 0>
__device__ __forceinline__ void sync_step_x4(const NetView &nv, const DrawView &dv, const unsigned char *blob, const Col &st,
                                             const Col &tmp, Draw<PBN_DRAW_PHILOX> &d, u32 &ublk) {
    u32 acc = 0;
    for (int i = 0; i < nv.n; i += 4) {
        u32 x0, x1, x2, x3;
        philox4x32_10_rk(ublk++, d.c1, d.c2, d.c3, dv, x0, x1, x2, x3);
        acc |= node_next<NET, PBN_DRAW_PHILOX, TQ>(nv, blob, st, i, d, x0, true) << (i & 31);
        acc |= node_next<NET, PBN_DRAW_PHILOX, TQ>(nv, blob, st, i + 1, d, x1, true) << ((i + 1) & 31);
        acc |= node_next<NET, PBN_DRAW_PHILOX, TQ>(nv, blob, st, i + 2, d, x2, true) << ((i + 2) & 31);
        acc |= node_next<NET, PBN_DRAW_PHILOX, TQ>(nv, blob, st, i + 3, d, x3, true) << ((i + 3) & 31);
        if (((i + 3) & 31) == 31 || i + 4 >= nv.n) { tmp.set_word(i >> 5, acc); acc = 0; }
    }
    for (int w = 0; w < nv.w32; w++) st.set_word(w, tmp.word(w));
}

// ----------------------------------------------------------------------------------------------- cubes
// cubes: u32 [n_cubes][w32][2] = (care, value); a state matches iff (word & care) == value for every word.
__device__ __forceinline__ bool cube_match(const u32 *cubes, int c, const Col &st, int w32) {
    const u32 *p = cubes + (size_t)c * w32 * 2;
    // networks up to 256 nodes: branch-free (the step-until-attractor tail is latency-bound, branches cost more than the
    // compares); larger ones: early exit per word (a full-care cube of a 1024-node network fails in its first word)
    if (w32 <= 8) {
        bool ok = true;
        for (int w = 0; w < w32; w++) ok &= ((st.word(w) & p[2 * w]) == p[2 * w + 1]);
        return ok;
    }
    for (int w = 0; w < w32; w++)
        if ((st.word(w) & p[2 * w]) != p[2 * w + 1]) return false;
    return true;
}
__device__ __forceinline__ bool match_range(const u32 *cubes, int c0, int c1, const Col &st, int w32) {
    for (int c = c0; c < c1; c++)
        if (cube_match(cubes, c, st, w32)) return true;
    return false;
}
