/*
 * pbn_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, byte-per-node CPU restatement of the reference's hot path
 * (jakub-zarzycki2022/gym-PBN-stac, pure Python).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load it; the product
 * (gym-pbn-stac_b200/) never does and fails loudly without its CUDA library.
 *
 * Parity pinning: the reference's own tests hold no golden vectors for this path
 * (SURVEY.md §8c), so the oracle is pinned against outputs of the reference itself,
 * recorded in this container by oracle/make_golden.py under replayed draws and committed
 * under tests/golden/ (tests/test_oracle_golden.py).
 *
 * Two draw sources:
 *   REPLAY — consumes recorded CPython `random` / numpy.random draws in the reference's exact
 *            order and compares in float64 exactly as the reference does;
 *   PHILOX — Philox4x32-10 keyed by (seed; epoch, global env id), integer 31-bit thresholds.
 *            Same structure, integer arithmetic; this is the stream the CUDA kernels use, so
 *            GPU-vs-oracle comparisons in this mode are bit-exact at any size.
 *
 * Each function cites the reference file:line it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_TT 0   /* truth-table PBN / PBCN  (gym_PBN/envs/common) */
#define ORC_PRED 1 /* Bittner predictor graph (gym_PBN/envs/bittner/base.py) */

typedef struct {
    int32_t kind, n, first; /* first updatable node: 1 for PBN.step (common/pbn.py:90), 0 for Graph.step (base.py:308) */
    /* ORC_TT: node i reads inputs tt_in[tt_in_off[i] .. tt_in_off[i+1]) (ascending node index, first = MSB,
       common/node.py:31-32) and P(next=1) = tt_prob[tt_tab_off[i] + idx] */
    const int32_t *tt_in_off, *tt_in, *tt_tab_off;
    const double *tt_prob;
    /* ORC_PRED: node i owns predictors pr_off[i] .. pr_off[i+1); predictor q reads nodes pr_in[4q..4q+3]
       (3 inputs then the node itself, base.py:100-104), LUT bit (x0<<3|x1<<2|x2<<1|x3) = [X.A >= 0] (base.py:110-118),
       pr_cum[q] = cumulative COD (base.py:30-45), pr_codsum[i] = CODsum */
    const int32_t *pr_off, *pr_in;
    const uint16_t *pr_lut;
    const double *pr_cum, *pr_codsum;
    /* PHILOX mode: 31-bit thresholds filled by orc_fill_thresholds (caller-allocated, same shapes as tt_prob / pr_cum) */
    uint32_t *tt_thr, *pr_thr;
} OrcNet;

#define ORC_PHILOX 0
#define ORC_REPLAY 1

typedef struct {
    int32_t mode;
    uint32_t epoch;
    uint64_t seed;
    const int32_t *ints; /* replay: row e = ints + e*int_stride */
    const double *dbls;
    int64_t int_stride, dbl_stride;
    int64_t *used; /* optional out [B][2]: ints/doubles (replay) or u32 draws/0 (philox) consumed */
} OrcDraws;

/* env kinds — one per reference env class on the hot path */
#define ORC_ENV_PBN 0      /* PBNEnv              pbn_env.py:125-188 */
#define ORC_ENV_PBCN 1     /* PBCNEnv             pbcn_env.py:52-80 */
#define ORC_ENV_TARGET 2   /* PBNTargetEnv        pbn_target.py:241-326 */
#define ORC_ENV_MULTI 3    /* PBNTargetMultiEnv   pbn_target_multi.py:119-225 */
#define ORC_ENV_PBN_SD 4   /* PBNSampledDataEnv   sampled_data.py:52-88 */
#define ORC_ENV_PBCN_SD 5  /* PBCNSampledDataEnv  sampled_data.py:139-189 */
#define ORC_ENV_PBN_ST 6   /* PBNSelfTriggeringEnv  self_triggering.py:56-93 */
#define ORC_ENV_PBCN_ST 7  /* PBCNSelfTriggeringEnv self_triggering.py:146-197 */

typedef struct {
    int32_t kind, horizon, max_inner;
    int32_t force;        /* PBNTargetEnv.step(force=True): exactly one update, pbn_target.py:270 */
    int32_t dedup;        /* multi: tensor actions are unique()'d, lists are not, pbn_target_multi.py:120-121 */
    int32_t control_write;/* PBCN: 0 = reference (apply_control has no effect on dynamics, Q4); 1 = write control into state[0:M] */
    int32_t n_control;
    int32_t successful_reward, wrong_attractor_cost; /* PBCNEnv._get_reward, pbcn_env.py:52-65 */
    /* cubes: attractor a owns cubes att_off[a] .. att_off[a+1); cube c = cube[c*n .. ), values 0/1/2('*') */
    int32_t n_att;
    const int32_t *att_off;
    const int8_t *cube;
    /* PBN/PBCN target set = full states, tgt_off = first cube of the target list, n_tgt entries (pbn_env.py:55-59) */
    int32_t tgt_first, n_tgt;
    /* self-triggering envs: gamma_pow[i] = gamma**i as Python computes it, i < n_gamma; max_interval = T (0 = no cap) */
    const double *gamma_pow;
    int32_t n_gamma, max_interval;
} OrcEnv;

/* ------------------------------------------------------------------------------------ Philox4x32-10 */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; r++) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof c);
}

typedef struct {
    int mode;
    uint32_t key[2], ctr[4], buf[4];
    int have;
    const int32_t *ip; const double *dp;
    int64_t ni, nd;
} Dr;

static void dr_init(Dr *d, const OrcDraws *s, int64_t local, int64_t env_id) {
    d->mode = s->mode; d->have = 0; d->ni = d->nd = 0;
    d->key[0] = (uint32_t)s->seed; d->key[1] = (uint32_t)(s->seed >> 32);
    d->ctr[0] = 0; d->ctr[1] = s->epoch; d->ctr[2] = (uint32_t)env_id; d->ctr[3] = (uint32_t)((uint64_t)env_id >> 32);
    d->ip = s->ints ? s->ints + local * s->int_stride : NULL;
    d->dp = s->dbls ? s->dbls + local * s->dbl_stride : NULL;
}
static void dr_done(const Dr *d, const OrcDraws *s, int64_t local) {
    if (s->used) { s->used[2 * local] = d->ni; s->used[2 * local + 1] = d->nd; }
}
static inline uint32_t dr_u32(Dr *d) { /* philox only: x,y,z,w of block 0, then block 1, ... */
    if (!d->have) { orc_philox4x32_10(d->ctr, d->key, d->buf); d->ctr[0]++; d->have = 4; }
    d->ni++;
    return d->buf[4 - d->have--];
}
/* uniform integer in [lo, lo+n): random.randint(lo, lo+n-1) */
static inline int dr_randint(Dr *d, int lo, int n) {
    if (d->mode == ORC_REPLAY) { d->ni++; return *d->ip++; }
    return lo + (int)(((uint64_t)dr_u32(d) * (uint32_t)n) >> 32);
}
static inline double dr_dbl(Dr *d) { d->nd++; return *d->dp++; }
/* '*' -> randint(0,1) of a reset: replay = one recorded draw each; Philox = the bits of the stream's words, most significant
   first, 32 wildcards per word */
typedef struct { uint32_t buf; int n; } WildBits;
static inline int dr_wild(Dr *d, WildBits *wb) {
    if (d->mode == ORC_REPLAY) return dr_randint(d, 0, 2);
    if (wb->n == 0) { wb->buf = dr_u32(d); wb->n = 32; }
    wb->n--;
    int v = (int)(wb->buf >> 31);
    wb->buf <<= 1;
    return v;
}
static inline uint32_t thr31(double p) { /* smallest T with (r31 < T) <=> (r31 / 2^31 < p) */
    double t = ceil(p * 2147483648.0);
    if (!(t > 0)) return 0;
    if (t > 2147483648.0) t = 2147483648.0;
    return (uint32_t)t;
}

/* next = (r31 < T): T = ceil(P * 2^31); predictor k chosen iff r31 < ceil(cum_k / CODsum * 2^31) and no earlier one was */
void orc_fill_thresholds(OrcNet *net) {
    if (net->kind == ORC_TT) {
        for (int q = 0; q < net->tt_tab_off[net->n]; q++) net->tt_thr[q] = thr31(net->tt_prob[q]);
    } else {
        for (int i = 0; i < net->n; i++)
            for (int q = net->pr_off[i]; q < net->pr_off[i + 1]; q++) net->pr_thr[q] = thr31(net->pr_cum[q] / net->pr_codsum[i]);
    }
}

/* geometric gap, PHILOX mode only: number of failures before the next success of a Bernoulli(p) process.
   u = ((r>>9)+0.5)/2^23 in (0,1);  G = trunc(log2(u) * inv), inv = 1/log2(1-p).
   log2 is a fixed polynomial evaluated with IEEE single fma only, so CPU and GPU agree bit for bit. */
static inline float orc_log2f_poly(float x) { /* x > 0, normal */
    uint32_t b; memcpy(&b, &x, 4);
    int e = (int)(b >> 23) - 127;
    b = (b & 0x007FFFFFu) | 0x3F800000u;
    float m; memcpy(&m, &b, 4);
    if (m > 1.41421356f) { m *= 0.5f; e += 1; }
    float t = m - 1.0f; /* in [-0.2929, 0.4142] */
    /* log2(1+t) = t * P(t); P = degree-7 Chebyshev-node interpolant on [1/sqrt2-1, sqrt2-1], |err| < 1.3e-7 */
    float p = -1.427597404e-01f;
    p = fmaf(p, t, 2.326525748e-01f);
    p = fmaf(p, t, -2.492718250e-01f);
    p = fmaf(p, t, 2.872888744e-01f);
    p = fmaf(p, t, -3.602251709e-01f);
    p = fmaf(p, t, 4.809167087e-01f);
    p = fmaf(p, t, -7.213529348e-01f);
    p = fmaf(p, t, 1.442695022e+00f);
    return fmaf(p, t, (float)e);
}
uint32_t orc_geom(uint32_t r, float inv) {
    float u = ((float)(r >> 9) + 0.5f) * (1.0f / 8388608.0f);
    float g = orc_log2f_poly(u) * inv;
    if (!(g < 33554432.0f)) return 33554432u; /* capped at 2^25: the 64 summed gaps of a round fit 32 bits */
    return (uint32_t)g;
}
float orc_geom_inv(double p) { /* host helper shared by tests: 1/log2(1-p) as float */
    if (p <= 0) return 1.0f;  /* sentinel (> 0): flips disabled; every valid value is <= 0 */
    if (p >= 1) return 0.0f;
    return (float)(1.0 / log2(1.0 - p));
}

/* ------------------------------------------------------------------------------------ node updates */
/* common/node.py:31-38 + common/pbn.py:88-92 (PBN.step) / common/pbcn.py:51-66 (PBCN.step) */
static inline int tt_next_value(const OrcNet *net, const uint8_t *st, int i, Dr *d) {
    int idx = 0;
    for (int q = net->tt_in_off[i]; q < net->tt_in_off[i + 1]; q++) idx = (idx << 1) | st[net->tt_in[q]];
    double p = net->tt_prob[net->tt_tab_off[i] + idx];
    if (d->mode == ORC_REPLAY) return dr_dbl(d) < p; /* u < p, node.py:37-38 */
    return (dr_u32(d) >> 1) < net->tt_thr[net->tt_tab_off[i] + idx];
}
/* bittner/base.py:89-119 Node.Predstep */
static inline int pred_next_value(const OrcNet *net, const uint8_t *st, int i, Dr *d) {
    int q0 = net->pr_off[i], q1 = net->pr_off[i + 1], q = q1 - 1; /* falls through to the LAST predictor, base.py:95-97 */
    double S = net->pr_codsum[i];
    if (d->mode == ORC_REPLAY) {
        double r = dr_dbl(d) * S; /* base.py:94 */
        for (int k = q0; k < q1; k++) if (net->pr_cum[k] > r) { q = k; break; }
    } else {
        uint32_t r = dr_u32(d) >> 1;
        for (int k = q0; k < q1 - 1; k++) if (r < net->pr_thr[k]) { q = k; break; }
    }
    const int32_t *in = net->pr_in + 4 * q;
    int idx = (st[in[0]] << 3) | (st[in[1]] << 2) | (st[in[2]] << 1) | st[in[3]];
    return (net->pr_lut[q] >> idx) & 1;
}
/* one asynchronous update: PBN.step common/pbn.py:88-92, Graph.step base.py:306-312 */
static inline void micro_step(const OrcNet *net, uint8_t *st, Dr *d) {
    int i = dr_randint(d, net->first, net->n - net->first);
    st[i] = (uint8_t)(net->kind == ORC_TT ? tt_next_value(net, st, i, d) : pred_next_value(net, st, i, d));
}
/* Graph.synch_step with perturbations off, base.py:300-303: every node from the OLD state, node order */
static inline void sync_step(const OrcNet *net, uint8_t *st, uint8_t *tmp, Dr *d) {
    for (int i = 0; i < net->n; i++)
        tmp[i] = (uint8_t)(net->kind == ORC_TT ? tt_next_value(net, st, i, d) : pred_next_value(net, st, i, d));
    memcpy(st, tmp, (size_t)net->n);
}

/* K1: `steps` updates for B envs; state is uint8 [B][n] */
int orc_rollout(const OrcNet *net, uint8_t *state, int64_t B, int64_t env0, int64_t steps, int sync, const OrcDraws *dr) {
#pragma omp parallel for schedule(static) if (B >= 256)
    for (int64_t e = 0; e < B; e++) {
        Dr d; dr_init(&d, dr, e, env0 + e);
        uint8_t *st = state + e * net->n;
        uint8_t *tmp = sync ? (uint8_t *)malloc((size_t)net->n) : NULL;
        for (int64_t t = 0; t < steps; t++) { if (sync) sync_step(net, st, tmp, &d); else micro_step(net, st, &d); }
        free(tmp);
        dr_done(&d, dr, e);
    }
    return 0;
}

/* K1, bit-sliced synchronous mode (PHILOX only, predictor graphs with at most 5 predictors per node).
   Same law as Graph.synch_step (base.py:300-303): every node of every env is redrawn from the OLD state, predictor
   chosen by cumulative COD weight.  Restated for groups of 32 consecutive global env ids handled as ONE 32-bit word
   per node (bit b = env 32*G + b):
     * the predictor choice of the 32 envs for node i at step t compares 32 independent 31-bit uniforms with the node's
       thresholds bit-serially, most significant bit first: level l draws one word R (bit b = bit 30-l of env b's
       uniform) from Philox4x32-10 with counter (l>>2, epoch, G, t<<12 | i), word l&3; for threshold T with bit 30-l set,
       envs still undecided whose bit is 0 are decided "r < T", those with bit 1 stay undecided; with the bit clear,
       undecided envs with bit 1 are decided "r >= T".  Levels stop as soon as no env is undecided for any threshold.
     * predictor j is chosen where r >= T_{j-1} and r < T_j; the node's next word is OR_j (chosen_j & f_j(old words)). */
int orc_rollout_sync_sliced(const OrcNet *net, uint8_t *state, int64_t B, int64_t env0, int64_t steps, const OrcDraws *dr) {
    const int n = net->n;
    if (net->kind != ORC_PRED || dr->mode != ORC_PHILOX || (env0 & 31)) return 1;
    for (int i = 0; i < n; i++) if (net->pr_off[i + 1] - net->pr_off[i] > 5) return 2;
    const int64_t groups = (B + 31) / 32;
    const uint32_t key[2] = {(uint32_t)dr->seed, (uint32_t)(dr->seed >> 32)};
#pragma omp parallel for schedule(static) if (B >= 256)
    for (int64_t gi = 0; gi < groups; gi++) {
        const int64_t gb = gi * 32;
        const uint64_t G = (uint64_t)(env0 + gb) >> 5;
        uint32_t *cur = (uint32_t *)calloc((size_t)n, 4), *nxt = (uint32_t *)calloc((size_t)n, 4);
        for (int i = 0; i < n; i++)
            for (int b = 0; b < 32 && gb + b < B; b++) cur[i] |= (uint32_t)state[(gb + b) * n + i] << b;
        for (int64_t t = 0; t < steps; t++) {
            for (int i = 0; i < n; i++) {
                const int q0 = net->pr_off[i], f = net->pr_off[i + 1] - q0;
                uint32_t lt[4] = {0, 0, 0, 0}, und[4] = {0, 0, 0, 0}, T[4] = {0, 0, 0, 0};
                for (int k = 0; k < f - 1; k++) {
                    T[k] = net->pr_thr[q0 + k];
                    if (T[k] >= 0x80000000u) lt[k] = 0xFFFFFFFFu; else und[k] = 0xFFFFFFFFu;
                }
                uint32_t w[4] = {0, 0, 0, 0};
                for (int l = 0; l < 31 && (und[0] | und[1] | und[2] | und[3]); l++) {
                    if ((l & 3) == 0) {
                        const uint32_t ctr[4] = {(uint32_t)(l >> 2), dr->epoch, (uint32_t)G, (uint32_t)((t << 12) | i)};
                        orc_philox4x32_10(ctr, key, w);
                    }
                    const uint32_t R = w[l & 3];
                    for (int k = 0; k < f - 1; k++) {
                        if ((T[k] >> (30 - l)) & 1u) { lt[k] |= und[k] & ~R; und[k] &= R; }
                        else und[k] &= ~R;
                    }
                }
                uint32_t out = 0;
                for (int j = 0; j < f; j++) {
                    const uint32_t below = j < f - 1 ? lt[j] : 0xFFFFFFFFu;      /* r < T_j (the last predictor takes the rest) */
                    const uint32_t not_prev = j > 0 ? ~lt[j - 1] : 0xFFFFFFFFu;  /* r >= T_{j-1} */
                    const uint32_t chosen = below & not_prev;
                    const int32_t *in = net->pr_in + 4 * (q0 + j);
                    const uint32_t x0 = cur[in[0]], x1 = cur[in[1]], x2 = cur[in[2]], x3 = cur[in[3]];
                    uint32_t fj = 0;
                    for (int idx = 0; idx < 16; idx++)
                        if ((net->pr_lut[q0 + j] >> idx) & 1)
                            fj |= ((idx & 8) ? x0 : ~x0) & ((idx & 4) ? x1 : ~x1) & ((idx & 2) ? x2 : ~x2) & ((idx & 1) ? x3 : ~x3);
                    out |= chosen & fj;
                }
                nxt[i] = out;
            }
            uint32_t *tmp = cur; cur = nxt; nxt = tmp;
        }
        for (int i = 0; i < n; i++)
            for (int b = 0; b < 32 && gb + b < B; b++) state[(gb + b) * n + i] = (uint8_t)((cur[i] >> b) & 1u);
        free(cur); free(nxt);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ cube matching */
static inline int cube_match(const int8_t *cube, const uint8_t *st, int n) {
    for (int i = 0; i < n; i++) if (cube[i] != 2 && cube[i] != (int8_t)st[i]) return 0;
    return 1;
}
/* Bittner7.is_attracting_state pbn_target.py:562-574 (any cube of any attractor); the multi env's
   set lookup over expanded wildcards (pbn_target_multi.py:438-454,489-492) accepts the same states */
static inline int is_attracting(const OrcEnv *env, const uint8_t *st, int n) {
    if (env->n_att == 0) return 1; /* all-attracting fixture: every state attracting (PBNEnv.is_attracting_state ≡ True, pbn_env.py:19-21) */
    for (int c = 0; c < env->att_off[env->n_att]; c++) if (cube_match(env->cube + (int64_t)c * n, st, n)) return 1;
    return 0;
}
/* PBNTargetEnv.in_target pbn_target.py:289-301: any cube of the target attractor */
static inline int in_target_any(const OrcEnv *env, const uint8_t *st, int n, int a) {
    for (int c = env->att_off[a]; c < env->att_off[a + 1]; c++) if (cube_match(env->cube + (int64_t)c * n, st, n)) return 1;
    return 0;
}
/* PBNTargetMultiEnv.in_target pbn_target_multi.py:190-199: returns False at the FIRST mismatch of the FIRST cube (Q12) */
static inline int in_target_first(const OrcEnv *env, const uint8_t *st, int n, int a) {
    if (env->att_off[a] == env->att_off[a + 1]) return 0;
    return cube_match(env->cube + (int64_t)env->att_off[a] * n, st, n);
}
static inline int in_target_set(const OrcEnv *env, const uint8_t *st, int n) { /* tuple(obs) in self.target_nodes, pbn_env.py:168 */
    for (int c = env->tgt_first; c < env->tgt_first + env->n_tgt; c++) if (cube_match(env->cube + (int64_t)c * n, st, n)) return 1;
    return 0;
}
/* PBCNEnv._get_reward pbcn_env.py:52-65 */
static inline int pbcn_reward(const OrcEnv *env, const uint8_t *st, int n, int *term) {
    if (in_target_set(env, st, n)) { *term = 1; return env->successful_reward; }
    int m = 0;
    for (int a = 0; a < env->n_att; a++) m += in_target_any(env, st, n, a);
    *term = 0;
    return -env->wrong_attractor_cost * m;
}

/* K2: one env.step for B envs.
   actions: int32 [B][K].  PBN/PBCN/TARGET: K=1.  MULTI: K actions (value <0 = absent slot).
   PBN_SD: K=2 (primitive action, interval).  PBCN_SD: K = 1 + n_control (interval, control bits).
   outputs: reward int32[B], term/trunc uint8[B], inner int32[B] (micro-steps executed), obs uint8[B][n]
   (obs differs from state only for MULTI's pre-update capture, Q11). */
static int env_step_impl(const OrcNet *net, const OrcEnv *env, uint8_t *state, int32_t *n_steps, const int32_t *target_att,
                         const int32_t *actions, int K, uint8_t *obs, int32_t *reward, uint8_t *term, uint8_t *trunc,
                         int32_t *inner, int64_t B, int64_t env0, const OrcDraws *dr, double *reward_f64) {
    const int n = net->n;
#pragma omp parallel for schedule(dynamic, 64) if (B >= 256)
    for (int64_t e = 0; e < B; e++) {
        Dr d; dr_init(&d, dr, e, env0 + e);
        uint8_t *st = state + e * n, *ob = obs + e * n;
        const int32_t *act = actions + e * K;
        int rew = 0, tm = 0, tr = 0, in = 0;
        switch (env->kind) {
        case ORC_ENV_PBN: { /* pbn_env.py:141-154, reward :171-183 */
            int a = act[0];
            if (a != 0) st[a] ^= 1; /* flips index `action` itself, Q3 */
            micro_step(net, st, &d); in = 1; /* is_attracting_state ≡ True: the while never iterates */
            if (in_target_set(env, st, n)) { rew = 20; tm = 1; } else rew = -4 - (a != 0);
            memcpy(ob, st, (size_t)n);
        } break;
        case ORC_ENV_PBCN: { /* pbcn_env.py:67-80 */
            int a = act[0];
            if (a != 0) st[a] ^= 1;
            micro_step(net, st, &d); in = 1;
            rew = pbcn_reward(env, st, n, &tm);
            memcpy(ob, st, (size_t)n);
        } break;
        case ORC_ENV_TARGET: { /* pbn_target.py:261-280, reward :303-326 */
            int a = act[0];
            n_steps[e] += 1;
            if (a != 0) st[a - 1] ^= 1;
            micro_step(net, st, &d); in = 1;
            while (!env->force && !is_attracting(env, st, n) && in < env->max_inner) { micro_step(net, st, &d); in++; }
            if (in_target_any(env, st, n, target_att[e])) { rew = 20; tm = 1; } else rew = -5;
            tr = (n_steps[e] == env->horizon);
            memcpy(ob, st, (size_t)n);
        } break;
        case ORC_ENV_MULTI: { /* pbn_target_multi.py:119-154, reward :201-225 */
            int cnt = 0;
            n_steps[e] += 1;
            for (int k = 0; k < K; k++) {
                int a = act[k];
                if (a < 0) continue;
                if (env->dedup) { int dup = 0; for (int j = 0; j < k; j++) dup |= (act[j] == a); if (dup) continue; }
                cnt++;
                if (a != 0) st[a - 1] ^= 1;
            }
            memcpy(ob, st, (size_t)n);  /* observation captured BEFORE the update, :133 */
            micro_step(net, st, &d); in = 1;
            while (!is_attracting(env, ob, n) && in < env->max_inner) { micro_step(net, st, &d); memcpy(ob, st, (size_t)n); in++; }
            if (in_target_first(env, ob, n, target_att[e])) { rew = 1000; tm = 1; }
            rew -= cnt;
            tr = (n_steps[e] == env->horizon);
        } break;
        case ORC_ENV_PBN_SD: { /* sampled_data.py:52-88 */
            int a = act[0], interval = act[1];
            for (int i = 0; i < interval; i++) {
                if (a != 0) st[a - 1] ^= 1;
                micro_step(net, st, &d); in++;
                if (in_target_set(env, st, n)) { rew += 20; tm = 1; } else { rew += -4 - (a != 0); tm = 0; }
            }
            memcpy(ob, st, (size_t)n);
        } break;
        case ORC_ENV_PBCN_SD: { /* sampled_data.py:139-189 */
            int interval = act[0], tstep = -1;
            for (int i = 0; i < interval; i++) {
                if (env->control_write) for (int c = 0; c < env->n_control; c++) st[c] = (uint8_t)(act[1 + c] != 0);
                micro_step(net, st, &d); in++;
                int r = pbcn_reward(env, st, n, &tm) - 1; /* time_step_cost = 1 */
                if (tstep >= 0) r -= env->successful_reward; /* overshoot penalty */
                else if (tm) tstep = i;
                rew += r;
            }
            memcpy(ob, st, (size_t)n);
        } break;
        case ORC_ENV_PBN_ST:    /* self_triggering.py:56-93: (action, prob 1..10) */
        case ORC_ENV_PBCN_ST: { /* self_triggering.py:146-197: (prob 1..10, control bits...) */
            const int pbcn = env->kind == ORC_ENV_PBCN_ST;
            const int a = pbcn ? 0 : act[0];
            const double prob = (double)act[pbcn ? 0 : 1] / 10.0; /* prob /= 10 */
            const uint32_t stop_thr = (uint32_t)ceil(prob * 2147483648.0);
            double total = 0.0;
            int i = 0, end = 0;
            while (!end) {
                int r;
                if (!pbcn) {
                    if (a != 0) st[a - 1] ^= 1;
                    micro_step(net, st, &d);
                    if (in_target_set(env, st, n)) { r = 20; tm = 1; } else { r = -4 - (a != 0); tm = 0; } /* PBNEnv._get_reward */
                } else {
                    if (env->control_write) for (int c = 0; c < env->n_control; c++) st[c] = (uint8_t)(act[1 + c] != 0);
                    micro_step(net, st, &d);
                    r = pbcn_reward(env, st, n, &tm) - 1; /* time step cost */
                }
                total += env->gamma_pow[i < env->n_gamma ? i : env->n_gamma - 1] * (double)r; /* total += gamma**i * reward */
                i++;
                int stop = d.mode == ORC_REPLAY ? (dr_dbl(&d) <= prob) : ((dr_u32(&d) >> 1) < stop_thr);
                end = stop || i == env->max_interval; /* random.uniform(0, 1) <= prob or i == T */
            }
            in = i;
            rew = (int)total;
            if (reward_f64) reward_f64[e] = total;
            memcpy(ob, st, (size_t)n);
        } break;
        default: break;
        }
        reward[e] = rew; term[e] = (uint8_t)tm; trunc[e] = (uint8_t)tr; inner[e] = in;
        dr_done(&d, dr, e);
    }
    return 0;
}
int orc_env_step(const OrcNet *net, const OrcEnv *env, uint8_t *state, int32_t *n_steps, const int32_t *target_att,
                 const int32_t *actions, int K, uint8_t *obs, int32_t *reward, uint8_t *term, uint8_t *trunc,
                 int32_t *inner, int64_t B, int64_t env0, const OrcDraws *dr) {
    return env_step_impl(net, env, state, n_steps, target_att, actions, K, obs, reward, term, trunc, inner, B, env0, dr, NULL);
}
int orc_env_step_f64(const OrcNet *net, const OrcEnv *env, uint8_t *state, int32_t *n_steps, const int32_t *target_att,
                     const int32_t *actions, int K, uint8_t *obs, int32_t *reward, double *reward_f64, uint8_t *term,
                     uint8_t *trunc, int32_t *inner, int64_t B, int64_t env0, const OrcDraws *dr) {
    return env_step_impl(net, env, state, n_steps, target_att, actions, K, obs, reward, term, trunc, inner, B, env0, dr, reward_f64);
}

/* reset for the envs selected by mask (NULL = all).
   TARGET: pbn_target.py:328-352 — sample(all_attractors, 2) -> choice(state cube), choice(target cube)
           -> per position: randint(0,1) for a '*' in state, then for a '*' in target.
   MULTI : pbn_target_multi.py:227-259 — state attractor = first, target = last (Q14), rest identical.
   PBN/PBCN family: pbn_env.py:190-213 — attractor with <= 10 states, uniform state from it, PBN.reset forces
           state[0]=0 (common/pbn.py:77).  The reference's discarded first choice() is not drawn in PHILOX mode;
           in REPLAY mode the recorded indices are (attractor index tries..., state index). */
/* ---- curriculum of PBNTargetMultiEnv (pbn_target_multi.py:159-181, 232-235) ----------------------------------------------
   sample_pair: np.random.choice(range(A), size=2, replace=False, p=probabilities) restated (numpy legacy RandomState.choice,
   replace=False with p): draw size - n_uniq uniforms, searchsorted(cdf, x, side='right') on the normalised cumulative sum,
   keep the first occurrence of each value, zero the mass of what was found, repeat until two distinct ids are there.
   u: the uniforms in the order the generator hands them out; returns how many were consumed. */
static int ss_right(const double *cdf, int A, double x) { int k = 0; while (k < A && cdf[k] <= x) k++; return k < A ? k : A - 1; }
int orc_sample_pair(const double *prob, int A, const double *u, int *ids) {
    double p[64], cdf[64];
    int used = 0, found = 0;
    if (A > 64) A = 64;
    for (int k = 0; k < A; k++) p[k] = prob[k];
    while (found < 2) {
        for (int k = 0; k < found; k++) p[ids[k]] = 0.0;
        double acc = 0.0;
        for (int k = 0; k < A; k++) { acc += p[k]; cdf[k] = acc; }
        for (int k = 0; k < A; k++) cdf[k] /= acc;
        int want = 2 - found, fresh[2], nf = 0;
        for (int k = 0; k < want; k++) fresh[k] = ss_right(cdf, A, u[used++]);
        for (int k = 0; k < want; k++) { /* np.unique(return_index) + sort: first occurrences, in draw order */
            int dup = 0;
            for (int j = 0; j < k; j++) dup |= fresh[j] == fresh[k];
            if (!dup) ids[found + nf++] = fresh[k];
        }
        found += nf;
    }
    return used;
}
/* rework_probas(episode_len), pbn_target_multi.py:159-181, on one probability row */
void orc_rework_probas(double *prob, int A, int s, int t, int episode_len) {
    const double eps = 1.0 * 1.0 / A, lo = 0.01 * 1.0 / A, hi = 0.5;
    if (episode_len < 20) {
        prob[s] -= eps; prob[t] -= eps;
        prob[s] = prob[s] > lo ? prob[s] : lo;
        prob[t] = prob[t] > lo ? prob[t] : lo;
    }
    if (episode_len >= 99) {
        prob[s] += eps; prob[t] += eps;
        prob[s] = prob[s] < hi ? prob[s] : hi;
        prob[t] = prob[t] < hi ? prob[t] : hi;
    }
    for (int k = 0; k < A; k++) prob[k] = lo > prob[k] ? lo : prob[k];
    /* Python's sum() over floats: CPython >= 3.12 (3.12.3 here) adds with Neumaier compensation (Python/bltinmodule.c,
       cs_add) and folds the compensation in at the end */
    double sum = 0.0, comp = 0.0;
    for (int k = 0; k < A; k++) {
        const double x = prob[k], t = sum + x;
        comp += fabs(sum) >= fabs(x) ? (sum - t) + x : (x - t) + sum;
        sum = t;
    }
    if (comp != 0.0 && isfinite(comp)) sum += comp;
    for (int k = 0; k < A; k++) prob[k] /= sum;
}

/* MULTI reset with the pair drawn from the env's own probability row (Philox: u = (word + 0.5) * 2^-32 per uniform, drawn
   before everything else as the reference does, :232); sample_pair = 0 keeps first -> last for the attractors (Q14) and only
   records the ids, which is what the reference does. */
int orc_env_reset_cur(const OrcNet *net, const OrcEnv *env, uint8_t *state, int32_t *n_steps, int32_t *target_att,
                      uint8_t *target_state, const uint8_t *mask, double *prob, int32_t *pair_ids, int sample_pair, int64_t B,
                      int64_t env0, const OrcDraws *dr) {
    const int n = net->n, A = env->n_att;
    if (env->kind != ORC_ENV_MULTI || dr->mode != ORC_PHILOX || A < 2 || A > 64) return 1;
    for (int64_t e = 0; e < B; e++) {
        if (mask && !mask[e]) continue;
        Dr d; dr_init(&d, dr, e, env0 + e);
        double u[3];
        int ids[2];
        for (int k = 0; k < 3; k++) u[k] = ((double)dr_u32(&d) + 0.5) * (1.0 / 4294967296.0); /* three words, always */
        orc_sample_pair(prob + e * A, A, u, ids);
        pair_ids[2 * e] = ids[0]; pair_ids[2 * e + 1] = ids[1];
        const int a = sample_pair ? ids[0] : 0, b = sample_pair ? ids[1] : A - 1;
        int cs = env->att_off[a] + dr_randint(&d, 0, env->att_off[a + 1] - env->att_off[a]);
        int ct = env->att_off[b] + dr_randint(&d, 0, env->att_off[b + 1] - env->att_off[b]);
        const int8_t *s = env->cube + (int64_t)cs * n, *t = env->cube + (int64_t)ct * n;
        uint8_t *st = state + e * n;
        WildBits wb = {0, 0};
        for (int i = 0; i < n; i++) {
            st[i] = (uint8_t)(s[i] == 2 ? dr_wild(&d, &wb) : s[i]);
            uint8_t tv = (uint8_t)(t[i] == 2 ? dr_wild(&d, &wb) : t[i]);
            if (target_state) target_state[e * n + i] = tv;
        }
        target_att[e] = b;
        n_steps[e] = 0;
        dr_done(&d, dr, e);
    }
    return 0;
}

int orc_env_reset(const OrcNet *net, const OrcEnv *env, uint8_t *state, int32_t *n_steps, int32_t *target_att,
                  uint8_t *target_state, const uint8_t *mask, int64_t B, int64_t env0, const OrcDraws *dr) {
    const int n = net->n;
    for (int64_t e = 0; e < B; e++) {
        if (mask && !mask[e]) continue;
        Dr d; dr_init(&d, dr, e, env0 + e);
        uint8_t *st = state + e * n;
        if (env->kind == ORC_ENV_TARGET || env->kind == ORC_ENV_MULTI) {
            int A = env->n_att, a, b;
            if (env->kind == ORC_ENV_TARGET) {
                if (d.mode == ORC_REPLAY) { a = dr_randint(&d, 0, A); b = dr_randint(&d, 0, A); }
                else { a = dr_randint(&d, 0, A); b = dr_randint(&d, 0, A - 1); if (b >= a) b++; }
            } else { a = 0; b = A - 1; }
            int cs = env->att_off[a] + dr_randint(&d, 0, env->att_off[a + 1] - env->att_off[a]);
            int ct = env->att_off[b] + dr_randint(&d, 0, env->att_off[b + 1] - env->att_off[b]);
            const int8_t *s = env->cube + (int64_t)cs * n, *t = env->cube + (int64_t)ct * n;
            WildBits wb = {0, 0};
            for (int i = 0; i < n; i++) {
                st[i] = (uint8_t)(s[i] == 2 ? dr_wild(&d, &wb) : s[i]);
                uint8_t tv = (uint8_t)(t[i] == 2 ? dr_wild(&d, &wb) : t[i]);
                if (target_state) target_state[e * n + i] = tv;
            }
            target_att[e] = b;
            n_steps[e] = 0;
        } else {
            int a;
            do { a = dr_randint(&d, 0, env->n_att); } while (env->att_off[a + 1] - env->att_off[a] > 10);
            int c = env->att_off[a] + dr_randint(&d, 0, env->att_off[a + 1] - env->att_off[a]);
            for (int i = 0; i < n; i++) st[i] = (uint8_t)env->cube[(int64_t)c * n + i];
            st[0] = 0;
            if (n_steps) n_steps[e] = 0;
        }
        dr_done(&d, dr, e);
    }
    return 0;
}

/* Graph.genRandState base.py:368-370: randint(0,1) per node */
int orc_rand_state(const OrcNet *net, uint8_t *state, int64_t B, int64_t env0, const OrcDraws *dr) {
    for (int64_t e = 0; e < B; e++) {
        Dr d; dr_init(&d, dr, e, env0 + e);
        for (int i = 0; i < net->n; i++) state[e * net->n + i] = (uint8_t)dr_randint(&d, 0, 2);
        dr_done(&d, dr, e);
    }
    return 0;
}

/* K3: utils/eval.py:76-103 _ssd_run for `chains` chains x `iters` iterations, model=None.
   Per iteration: hist[bucket(state)] += 1 (before stepping, :88-89); flip each node w.p. p (:92-95);
   env.step(0) (:96) = PBNTargetEnv.step: one update, then until attracting (cap max_inner).
   REPLAY: n doubles per chain per iteration, then the step's draws.
   PHILOX: the flips of each GROUP of 32 consecutive chain ids (global id >> 5) are one Bernoulli(p) renewal
   process over the interleaved index c = node*32 + lane, window = 32*n positions per iteration: in every round each
   of the 32 lanes draws TWO geometric gaps from consecutive words of its own PERTURBATION stream (Philox block indices
   from 2^31 on; the update stream keeps block indices from 0 and is consumed at two words per update), the 64 events of
   a round are ordered lane-major (lane l holds events 2l and 2l+1), event k sits at (last event) + sum of (1+gap) over
   events 0..k, and an event inside the window flips node c>>5 of chain c&31.
   That is the same law as n independent Bernoulli(p) draws per chain per iteration, at ~1 draw per chain.
   bucket = target-node bits MSB-first (pbn_target.py:383-391).  hist is uint64 [2^g], summed over chains. */
int orc_ssd(const OrcNet *net, const OrcEnv *env, uint8_t *state, int64_t chains, int64_t env0, int64_t iters,
            double p, const int32_t *tgt_nodes, int g, uint64_t *hist, const OrcDraws *dr) {
    const int n = net->n;
    const int64_t nb = (int64_t)1 << g;
    const float inv = orc_geom_inv(p);
    const int flips = !(inv > 0);
    const uint32_t W = (uint32_t)n * 32u, NONE = 0xFFFFFFFFu;
    if (dr->mode == ORC_PHILOX && flips && (env0 & 31)) return 1;
    int nthreads = 1;
#ifdef _OPENMP
    extern int omp_get_max_threads(void); extern int omp_get_thread_num(void);
    nthreads = omp_get_max_threads();
#endif
    uint64_t *priv = (uint64_t *)calloc((size_t)(nthreads * nb), sizeof(uint64_t));
    const int64_t groups = (chains + 31) / 32;
#pragma omp parallel for schedule(static) if (chains >= 256)
    for (int64_t gi = 0; gi < groups; gi++) {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        uint64_t *h = priv + (int64_t)tid * nb;
        const int64_t gb = gi * 32;
        Dr d[32], dp[32]; /* per chain: update stream, and perturbation stream (same key/counter words, block indices from 2^31) */
        uint32_t ev[64], last_p1 = 0;
        for (int l = 0; l < 32; l++) {
            dr_init(&d[l], dr, gb + l < chains ? gb + l : 0, env0 + gb + l);
            dp[l] = d[l];
            dp[l].ctr[0] = 0x80000000u;
            ev[2 * l] = ev[2 * l + 1] = NONE;
        }
        for (int64_t t = 0; t < iters; t++) {
            for (int l = 0; l < 32 && gb + l < chains; l++) {
                const uint8_t *st = state + (gb + l) * n;
                int64_t b = 0;
                for (int k = 0; k < g; k++) b = (b << 1) | st[tgt_nodes[k]];
                h[b]++;
            }
            if (dr->mode == ORC_REPLAY) {
                for (int l = 0; l < 32 && gb + l < chains; l++) {
                    uint8_t *st = state + (gb + l) * n;
                    for (int j = 0; j < n; j++) if (dr_dbl(&d[l]) < p) st[j] ^= 1; /* graph.flipNode(j), eval.py:92-95 */
                }
            } else if (flips) {
                for (;;) {
                    for (int k = 0; k < 64; k++)
                        if (ev[k] < W) {
                            uint32_t tl = ev[k] & 31u, bit = ev[k] >> 5;
                            if (gb + tl < chains) state[(gb + tl) * n + bit] ^= 1;
                            ev[k] = NONE;
                        }
                    if (last_p1 > W) break;
                    uint32_t pre = 0;
                    for (int k = 0; k < 64; k++) { pre += 1u + orc_geom(dr_u32(&dp[k >> 1]), inv); ev[k] = last_p1 - 1u + pre; }
                    last_p1 = ev[63] + 1u;
                }
                for (int k = 0; k < 64; k++) if (ev[k] != NONE) ev[k] -= W;
                last_p1 -= W;
            }
            for (int l = 0; l < 32 && gb + l < chains; l++) {
                uint8_t *st = state + (gb + l) * n;
                micro_step(net, st, &d[l]);
                int in = 1;
                while (env && !env->force && !is_attracting(env, st, n) && in < env->max_inner) { micro_step(net, st, &d[l]); in++; }
            }
        }
        for (int l = 0; l < 32 && gb + l < chains; l++) dr_done(&d[l], dr, gb + l);
    }
    for (int t = 0; t < nthreads; t++) for (int64_t b = 0; b < nb; b++) hist[b] += priv[t * nb + b];
    free(priv);
    return 0;
}

/* exact law of orc_geom: counts[k] = #{x in [0, 2^23) : G(x << 9) == k} for k < kmax (last bin collects the tail) */
void orc_geom_law(double p, int64_t *counts, int kmax) {
    const float inv = orc_geom_inv(p);
    for (uint32_t x = 0; x < (1u << 23); x++) {
        uint32_t g = orc_geom(x << 9, inv);
        counts[g < (uint32_t)kmax ? g : (uint32_t)kmax - 1]++;
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
