/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the MCWF trajectory path of NoisyCircuits.
 *
 * This file is a plain-C restatement of the reference algorithm (reference paths are relative to
 * /root/reference/src/NoisyCircuits/utils/custom/src/).  It is imported only by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.  The product
 * (noisycircuits_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every entry point below against
 * the reference's own extensions compiled from their sources (oracle/_ref, built by oracle/Makefile)
 * and tests/golden/ holds vectors generated from those extensions.
 *
 * What is restated (reference file:line):
 *   gates X/SX/RZ/RX/RY/H/P            QuantumGates.hpp:449-700
 *   gates CZ/RZZ/ECR/CX/SWAP           QuantumGates.hpp:726-942
 *   generic k-qubit unitary            QuantumGates.hpp:968-1064
 *   Kraus branch probabilities         QuantumGates.hpp:44-58, 205-227
 *   jump + renormalise                 QuantumGates.hpp:117-144, 360-419
 *   one trajectory                     Simulator.cpp:76-94, SimulatorMPI.cpp:40-59
 *   trajectory average                 Simulator.cpp:118-133
 *   measurement error                  MeasurementErrorApplicator.cpp:65-114
 *   RNG: std::mt19937_64 + std::discrete_distribution<> (libstdc++ 13 bits/random.tcc)
 *
 * `mask_fix`: the reference enumerates two-qubit quads with m2 = (1<<(q_max-1))-1
 * (QuantumGates.hpp:367,730,770,825,880,927) which is only right for |q1-q2| == 1.  mask_fix=0
 * reproduces that bit-for-bit; mask_fix=1 uses (1<<q_max)-1 (the mathematically intended quads).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } c128;

enum {
    OP_X = 0, OP_SX = 1, OP_RZ = 2, OP_RX = 3, OP_RY = 4, OP_H = 5, OP_P = 6,
    OP_CZ = 7, OP_RZZ = 8, OP_ECR = 9, OP_CX = 10, OP_SWAP = 11, OP_UNITARY = 12
};

static inline c128 cmul(c128 a, c128 b) { c128 r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static inline c128 cadd(c128 a, c128 b) { c128 r = { a.re + b.re, a.im + b.im }; return r; }
static inline c128 csub(c128 a, c128 b) { c128 r = { a.re - b.re, a.im - b.im }; return r; }
static inline double cnorm2(c128 a) { return a.re * a.re + a.im * a.im; }

/* ---------------------------------------------------------------- index maps ------------- */
/* pair index -> (i, j): insert a zero bit at position q (QuantumGates.hpp:48-49). */
static inline size_t pair_lo(size_t p, size_t stride) { return (p & (stride - 1)) | ((p & ~(stride - 1)) << 1); }

/* quad base: insert zero bits at q_min then q_max (QuantumGates.hpp:365-368 and siblings). */
typedef struct { size_t m1, m2, b1, b2; } quad_map;
static quad_map make_quad_map(int q1, int q2, int mask_fix) {
    int q_min = q1 < q2 ? q1 : q2, q_max = q1 > q2 ? q1 : q2;
    quad_map m;
    m.m1 = ((size_t)1 << q_min) - 1;
    m.m2 = mask_fix ? (((size_t)1 << q_max) - 1) : (((size_t)1 << (q_max - 1)) - 1);
    m.b1 = (size_t)1 << q1;
    m.b2 = (size_t)1 << q2;
    return m;
}
static inline size_t quad_base(size_t i, const quad_map *m) {
    size_t s = (i & m->m1) | ((i & ~m->m1) << 1);
    return (s & m->m2) | ((s & ~m->m2) << 1);
}

/* ---------------------------------------------------------------- gates ------------------- */
static void gate_1q(c128 *st, int n, int op, int q, double theta) {
    const size_t dim = (size_t)1 << n, stride = (size_t)1 << q;
    double c = 0.0, s = 0.0;
    if (op == OP_P) { c = cos(theta); s = sin(theta); }           /* QuantumGates.hpp:689-690: cos(theta), not theta/2 */
    else { c = cos(0.5 * theta); s = sin(0.5 * theta); }
    const c128 hp = { 0.5, 0.5 }, hm = { 0.5, -0.5 };
    const double r2 = 0.7071067811865475;                          /* QuantumGates.hpp:650 */
    for (size_t p = 0; p < (dim >> 1); ++p) {
        const size_t i = pair_lo(p, stride), j = i | stride;
        const c128 a = st[i], b = st[j];
        c128 x, y;
        switch (op) {
        case OP_X:  x = b; y = a; break;
        case OP_SX: x = cadd(cmul(hp, a), cmul(hm, b)); y = cadd(cmul(hm, a), cmul(hp, b)); break;
        case OP_RZ: x.re = c * a.re + s * a.im; x.im = c * a.im - s * a.re;
                    y.re = c * b.re - s * b.im; y.im = c * b.im + s * b.re; break;
        case OP_RX: x.re = c * a.re + s * b.im; x.im = c * a.im - s * b.re;
                    y.re = s * a.im + c * b.re; y.im = c * b.im - s * a.re; break;
        case OP_RY: x.re = c * a.re - s * b.re; x.im = c * a.im - s * b.im;
                    y.re = s * a.re + c * b.re; y.im = s * a.im + c * b.im; break;
        case OP_H:  x.re = r2 * (a.re + b.re); x.im = r2 * (a.im + b.im);
                    y.re = r2 * (a.re - b.re); y.im = r2 * (a.im - b.im); break;
        case OP_P:  x = a; y.re = b.re * c - b.im * s; y.im = b.re * s + b.im * c; break;
        default:    x = a; y = b; break;
        }
        st[i] = x; st[j] = y;
    }
}

static void gate_2q(c128 *st, int n, int op, int q1, int q2, double theta, int mask_fix) {
    const size_t iters = ((size_t)1 << n) >> 2;
    const quad_map m = make_quad_map(q1, q2, mask_fix);
    const double c = cos(0.5 * theta), s = sin(0.5 * theta);
    const double r2 = 0.7071067811865476;                          /* QuantumGates.hpp:829 */
    for (size_t i = 0; i < iters; ++i) {
        const size_t pos = quad_base(i, &m);
        const size_t i00 = pos, i01 = pos | m.b1, i10 = pos | m.b2, i11 = pos | m.b1 | m.b2;
        const c128 s00 = st[i00], s01 = st[i01], s10 = st[i10], s11 = st[i11];
        switch (op) {
        case OP_CZ:
            st[i11].re = -s11.re; st[i11].im = -s11.im; break;
        case OP_RZZ:
            st[i00].re = c * s00.re + s * s00.im; st[i00].im = c * s00.im - s * s00.re;
            st[i01].re = c * s01.re - s * s01.im; st[i01].im = c * s01.im + s * s01.re;
            st[i10].re = c * s10.re - s * s10.im; st[i10].im = c * s10.im + s * s10.re;
            st[i11].re = c * s11.re + s * s11.im; st[i11].im = c * s11.im - s * s11.re; break;
        case OP_ECR:
            /* rows: |00> <- (s01 + i s11)/r2 ... QuantumGates.hpp:843-846 */
            st[i00].re = r2 * s01.re - r2 * s11.im; st[i00].im = r2 * s01.im + r2 * s11.re;
            st[i01].re = r2 * s00.re + r2 * s10.im; st[i01].im = r2 * s00.im - r2 * s10.re;
            st[i10].re = r2 * s11.re - r2 * s01.im; st[i10].im = r2 * s11.im + r2 * s01.re;
            st[i11].re = r2 * s10.re + r2 * s00.im; st[i11].im = r2 * s10.im - r2 * s00.re; break;
        case OP_CX:
            st[i01] = s11; st[i11] = s01; break;
        case OP_SWAP:
            st[i01] = s10; st[i10] = s01; break;
        default: break;
        }
    }
}

/* Dense 2^k x 2^k unitary on `targets` (already in the order the C++ sees them, i.e. the caller
 * has reversed the Python list as Simulator.cpp:182 does): matrix bit b <-> targets[b]. */
void orc_apply_unitary(c128 *st, int n, const int *targets, int k, const c128 *U) {
    const size_t dim = (size_t)1 << n, inner = (size_t)1 << k, outer = dim >> k;
    size_t *offs = (size_t *)malloc(inner * sizeof(size_t));
    c128 *tmp = (c128 *)malloc(inner * sizeof(c128));
    c128 *out = (c128 *)malloc(inner * sizeof(c128));
    unsigned char is_t[64] = { 0 };
    for (int t = 0; t < k; ++t) is_t[targets[t]] = 1;
    for (size_t a = 0; a < inner; ++a) {
        size_t idx = 0;
        for (int b = 0; b < k; ++b) if (a & ((size_t)1 << b)) idx |= (size_t)1 << targets[b];
        offs[a] = idx;
    }
    for (size_t o = 0; o < outer; ++o) {
        size_t base = 0, src = o;
        for (int bit = 0; bit < n; ++bit) {
            if (is_t[bit]) continue;
            if (src & 1) base |= (size_t)1 << bit;
            src >>= 1;
        }
        for (size_t a = 0; a < inner; ++a) tmp[a] = st[base | offs[a]];
        for (size_t r = 0; r < inner; ++r) {
            c128 acc = { 0.0, 0.0 };
            for (size_t c = 0; c < inner; ++c) acc = cadd(acc, cmul(U[r * inner + c], tmp[c]));
            out[r] = acc;
        }
        for (size_t a = 0; a < inner; ++a) st[base | offs[a]] = out[a];
    }
    free(offs); free(tmp); free(out);
}

void orc_apply_gate(c128 *st, int n, int op, int q0, int q1, double theta, int mask_fix) {
    if (op <= OP_P) gate_1q(st, n, op, q0, theta);
    else if (op <= OP_SWAP) gate_2q(st, n, op, q0, q1, theta, mask_fix);
}

/* ---------------------------------------------------------------- noise ------------------- */
/* p = || K psi ||^2 for one 2x2 operator (QuantumGates.hpp:44-58). K is row-major. */
static double prob_1q(const c128 *st, int n, int q, const c128 *K) {
    const size_t dim = (size_t)1 << n, stride = (size_t)1 << q;
    double p = 0.0;
    for (size_t pr = 0; pr < (dim >> 1); ++pr) {
        const size_t i = pair_lo(pr, stride), j = i | stride;
        const c128 n0 = cadd(cmul(K[0], st[i]), cmul(K[1], st[j]));
        const c128 n1 = cadd(cmul(K[2], st[i]), cmul(K[3], st[j]));
        p += n0.re * n0.re + n0.im * n0.im + n1.re * n1.re + n1.im * n1.im;
    }
    return p;
}
static void apply_1q(c128 *st, int n, int q, const c128 *K) {
    const size_t dim = (size_t)1 << n, stride = (size_t)1 << q;
    for (size_t pr = 0; pr < (dim >> 1); ++pr) {
        const size_t i = pair_lo(pr, stride), j = i | stride;
        const c128 a = st[i], b = st[j];
        st[i] = cadd(cmul(K[0], a), cmul(K[1], b));
        st[j] = cadd(cmul(K[2], a), cmul(K[3], b));
    }
}
/* 4x4 operator; row/col index = bit(q1) + 2*bit(q2) (QuantumGates.hpp:209-219). */
static double prob_2q(const c128 *st, int n, const quad_map *m, const c128 *K) {
    const size_t iters = ((size_t)1 << n) >> 2;
    double p = 0.0;
    for (size_t i = 0; i < iters; ++i) {
        const size_t pos = quad_base(i, m);
        const c128 s[4] = { st[pos], st[pos | m->b1], st[pos | m->b2], st[pos | m->b1 | m->b2] };
        double acc = 0.0;
        for (int r = 0; r < 4; ++r) {
            c128 v = cmul(K[4 * r], s[0]);
            v = cadd(v, cmul(K[4 * r + 1], s[1]));
            v = cadd(v, cmul(K[4 * r + 2], s[2]));
            v = cadd(v, cmul(K[4 * r + 3], s[3]));
            acc += v.re * v.re + v.im * v.im;
        }
        p += acc;
    }
    return p;
}
static void apply_2q(c128 *st, int n, const quad_map *m, const c128 *K) {
    const size_t iters = ((size_t)1 << n) >> 2;
    for (size_t i = 0; i < iters; ++i) {
        const size_t pos = quad_base(i, m);
        const size_t idx[4] = { pos, pos | m->b1, pos | m->b2, pos | m->b1 | m->b2 };
        const c128 s[4] = { st[idx[0]], st[idx[1]], st[idx[2]], st[idx[3]] };
        for (int r = 0; r < 4; ++r) {
            c128 v = cmul(K[4 * r], s[0]);
            v = cadd(v, cmul(K[4 * r + 1], s[1]));
            v = cadd(v, cmul(K[4 * r + 2], s[2]));
            v = cadd(v, cmul(K[4 * r + 3], s[3]));
            st[idx[r]] = v;
        }
    }
}

/* std::discrete_distribution<> semantics (libstdc++ 13 bits/random.tcc:2657-2714):
 * weights are normalised by their sum, partial sums are formed left to right, the last partial sum
 * is forced to 1.0, and the draw is lower_bound(cp, u).  Fewer than two weights: outcome 0. */
int orc_discrete_pick(const double *w, int K, double u) {
    if (K < 2) return 0;
    double sum = 0.0;
    for (int k = 0; k < K; ++k) sum += w[k];
    double cp = 0.0;
    for (int k = 0; k < K - 1; ++k) {
        cp += w[k] / sum;
        if (!(cp < u)) return k;          /* first k with cp[k] >= u */
    }
    return K - 1;                          /* cp[K-1] := 1.0 >= u always */
}

/* One noise step (QuantumGates.hpp:117-144 / 360-419) with an injected uniform `u`.
 * dim2 = 2 (1q, qubit q0) or 4 (2q on q0,q1).  kraus: K matrices, row-major, dim2*dim2 each.
 * Writes the K branch probabilities to probs (if non-NULL); returns the chosen branch. */
int orc_noise_step(c128 *st, int n, int q0, int q1, const c128 *kraus, int K, int dim2, double u,
                   int mask_fix, double *probs) {
    double stackp[64];
    double *p = K <= 64 ? stackp : (double *)malloc((size_t)K * sizeof(double));
    quad_map m; memset(&m, 0, sizeof(m));
    if (dim2 == 4) m = make_quad_map(q0, q1, mask_fix);
    for (int k = 0; k < K; ++k)
        p[k] = dim2 == 2 ? prob_1q(st, n, q0, kraus + 4 * (size_t)k) : prob_2q(st, n, &m, kraus + 16 * (size_t)k);
    const int c = orc_discrete_pick(p, K, u);
    if (dim2 == 2) apply_1q(st, n, q0, kraus + 4 * (size_t)c); else apply_2q(st, n, &m, kraus + 16 * (size_t)c);
    const double scale = 1.0 / sqrt(p[c]);
    const size_t dim = (size_t)1 << n;
    for (size_t i = 0; i < dim; ++i) { st[i].re *= scale; st[i].im *= scale; }
    if (probs) memcpy(probs, p, (size_t)K * sizeof(double));
    if (p != stackp) free(p);
    return c;
}

/* ---------------------------------------------------------------- RNG --------------------- */
/* std::mt19937_64 (Matsumoto/Nishimura 64-bit MT, libstdc++ parameters). */
typedef struct { uint64_t mt[312]; int idx; } mt64;
void orc_mt64_seed(mt64 *g, uint64_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 312; ++i) g->mt[i] = 6364136223846793005ULL * (g->mt[i - 1] ^ (g->mt[i - 1] >> 62)) + (uint64_t)i;
    g->idx = 312;
}
uint64_t orc_mt64_next(mt64 *g) {
    if (g->idx >= 312) {
        const uint64_t UM = 0xFFFFFFFF80000000ULL, LM = 0x7FFFFFFFULL, A = 0xB5026F5AA96619E9ULL;
        for (int i = 0; i < 312; ++i) {
            uint64_t x = (g->mt[i] & UM) | (g->mt[(i + 1) % 312] & LM);
            g->mt[i] = g->mt[(i + 156) % 312] ^ (x >> 1) ^ ((x & 1) ? A : 0);
        }
        g->idx = 0;
    }
    uint64_t x = g->mt[g->idx++];
    x ^= (x >> 29) & 0x5555555555555555ULL;
    x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
    x ^= (x << 37) & 0xFFF7EEE000000000ULL;
    x ^= x >> 43;
    return x;
}
/* std::generate_canonical<double,53>(mt19937_64): one 64-bit draw / 2^64, clamped below 1
 * (bits/random.tcc:3349-3381). */
double orc_mt64_canonical(mt64 *g) {
    double r = (double)orc_mt64_next(g) / 18446744073709551616.0;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}
/* The uniform stream the reference consumes for seed `seed`: one draw per noisy instruction whose
 * Kraus list has >= 2 operators; instructions with 0/1 operators consume nothing.
 * out[g] is left at 0.5 for those. */
void orc_reference_uniforms(uint64_t seed, int G, const int *kraus_count, double *out) {
    mt64 g; orc_mt64_seed(&g, seed);
    for (int i = 0; i < G; ++i) out[i] = kraus_count[i] >= 2 ? orc_mt64_canonical(&g) : 0.5;
}

/* ---------------------------------------------------------------- trajectories ------------ */
/* Program layout shared by all entry points below:
 *   G instructions: op[g], q0[g], q1[g], theta[g]; chan[g] = channel id or -1 (no noise step);
 *   unitary gates: uni_off[g] = offset (in c128) into uni_pool, uni_k[g] = #targets,
 *                  uni_tq[8*g + b] = target for matrix bit b (already reversed);
 *   channels: ch_off[c] (in c128) into kraus_pool, ch_cnt[c] operators, ch_dim[c] = 2 or 4. */
typedef struct {
    int n, G;
    const int *op, *q0, *q1; const double *theta; const int *chan;
    const int64_t *uni_off; const int *uni_k; const int *uni_tq; const c128 *uni_pool;
    const int64_t *ch_off; const int *ch_cnt; const int *ch_dim; const c128 *kraus_pool;
} orc_program;

/* One trajectory from |0..0> with per-instruction uniforms (Simulator.cpp:76-94 with the
 * discrete_distribution draw replaced by u[g]).  choices/pc (optional, length G) receive the branch
 * index and its probability; state receives the final normalised state. */
void orc_trajectory(const orc_program *P, const double *u, int mask_fix, c128 *state, int *choices, double *pc) {
    const size_t dim = (size_t)1 << P->n;
    memset(state, 0, dim * sizeof(c128));
    state[0].re = 1.0;
    double *probs = NULL; int probs_cap = 0;
    for (int g = 0; g < P->G; ++g) {
        if (P->op[g] == OP_UNITARY)
            orc_apply_unitary(state, P->n, P->uni_tq + 8 * g, P->uni_k[g], P->uni_pool + P->uni_off[g]);
        else
            orc_apply_gate(state, P->n, P->op[g], P->q0[g], P->q1[g], P->theta[g], mask_fix);
        const int ch = P->chan[g];
        if (choices) choices[g] = -1;
        if (pc) pc[g] = 1.0;
        if (ch < 0) continue;
        const int K = P->ch_cnt[ch];
        if (K > probs_cap) { free(probs); probs = (double *)malloc((size_t)K * sizeof(double)); probs_cap = K; }
        const int c = orc_noise_step(state, P->n, P->q0[g], P->q1[g], P->kraus_pool + P->ch_off[ch], K, P->ch_dim[ch],
                                     u[g], mask_fix, probs);
        if (choices) choices[g] = c;
        if (pc) pc[g] = probs[c];
    }
    free(probs);
}

/* Mean of |psi|^2 over T trajectories (Simulator.cpp:118-133).
 *   uniforms != NULL: row t holds the G uniforms of trajectory t (replay mode);
 *   uniforms == NULL: reference seeding, trajectory t uses mt19937_64(seed0 + t)  (Simulator.cpp:124).
 * probs (length 2^n) is overwritten.  Accumulation is in trajectory order per thread and the
 * per-thread partials are summed in thread order, so the result is deterministic for a given
 * thread count. */
void orc_mcwf(const orc_program *P, int64_t T, const double *uniforms, uint64_t seed0, int mask_fix,
              int threads, double *probs) {
    const size_t dim = (size_t)1 << P->n;
    int *cnt = (int *)malloc((size_t)P->G * sizeof(int));
    for (int g = 0; g < P->G; ++g) cnt[g] = P->chan[g] >= 0 ? P->ch_cnt[P->chan[g]] : 0;
    if (threads < 1) threads = 1;
    double *partial = (double *)calloc((size_t)threads * dim, sizeof(double));
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        const int tid = 0, nt = 1;
#endif
        c128 *st = (c128 *)malloc(dim * sizeof(c128));
        double *u = (double *)malloc((size_t)P->G * sizeof(double));
        double *acc = partial + (size_t)tid * dim;
        for (int64_t t = tid; t < T; t += nt) {
            const double *ut = uniforms ? uniforms + (size_t)t * P->G : u;
            if (!uniforms) orc_reference_uniforms(seed0 + (uint64_t)t, P->G, cnt, u);
            orc_trajectory(P, ut, mask_fix, st, NULL, NULL);
            for (size_t i = 0; i < dim; ++i) acc[i] += cnorm2(st[i]);
        }
        free(st); free(u);
    }
    const double inv = 1.0 / (double)T;
    for (size_t i = 0; i < dim; ++i) {
        double s = 0.0;
        for (int t = 0; t < threads; ++t) s += partial[(size_t)t * dim + i];
        probs[i] = s * inv;
    }
    free(partial); free(cnt);
}

/* Noise-free run (Simulator.cpp:34-51): state receives the final state vector. */
void orc_pure_state(const orc_program *P, int mask_fix, c128 *state) {
    const size_t dim = (size_t)1 << P->n;
    memset(state, 0, dim * sizeof(c128));
    state[0].re = 1.0;
    for (int g = 0; g < P->G; ++g) {
        if (P->op[g] == OP_UNITARY)
            orc_apply_unitary(state, P->n, P->uni_tq + 8 * g, P->uni_k[g], P->uni_pool + P->uni_off[g]);
        else
            orc_apply_gate(state, P->n, P->op[g], P->q0[g], P->q1[g], P->theta[g], mask_fix);
    }
}

/* ---------------------------------------------------------------- measurement error ------- */
/* MeasurementErrorApplicator.cpp:65-81, 101-114: for the k-th listed qubit, the k-th matrix
 * (row-major 2x2, positional) is applied on probability pairs at stride 1<<qubits[k], in a vector of
 * 2^num_qubits entries. */
void orc_measurement_error(double *probs, int num_qubits, const double *mats, const int *qubits, int nq) {
    const size_t dim = (size_t)1 << num_qubits;
    for (int k = 0; k < nq; ++k) {
        const double m00 = mats[4 * k], m01 = mats[4 * k + 1], m10 = mats[4 * k + 2], m11 = mats[4 * k + 3];
        const size_t stride = (size_t)1 << qubits[k];
        for (size_t p = 0; p < (dim >> 1); ++p) {
            const size_t i = pair_lo(p, stride), j = i | stride;
            const double p0 = probs[i], p1 = probs[j];
            probs[i] = m00 * p0 + m01 * p1;
            probs[j] = m10 * p0 + m11 * p1;
        }
    }
}
