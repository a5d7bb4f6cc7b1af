// Shared host/device definitions of the B200 noisy-circuit engine: packed program format, Philox draws, the
// shared-memory tile layout and the per-thread executors.  Everything marked NCB_HD compiles both as CUDA device
// code (the product) and as plain C++ (tests/emu: a thread-by-thread emulation used only to test the host planner
// and the index maps on machines without a GPU).
#pragma once
#if defined(__CUDACC_RTC__)
// in-process compilation of the generated kernels (NVRTC): no host headers
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <math.h>
#include <stdint.h>
#endif

#if defined(__CUDACC__)
#define NCB_HD __host__ __device__ __forceinline__
#define NCB_D __device__ __forceinline__
#if defined(__CUDA_ARCH__)
// keeps the compiler from hoisting a whole operator matrix into registers across amplitude groups
#define NCB_NO_HOIST() asm volatile("" ::: "memory")
#define NCB_UNREACHABLE() __builtin_unreachable()
#endif
#else
#define NCB_HD inline
#define NCB_D inline
#endif

#ifndef NCB_NO_HOIST
#define NCB_NO_HOIST() ((void)0)
#endif
#ifndef NCB_UNREACHABLE
#define NCB_UNREACHABLE() ((void)0)
#endif

namespace ncb {

// Load of program data (ops, matrices).  SMEM = true: the pointer is known to address shared memory (the tile kernel's
// staged segment blob), so explicit ld.shared is emitted instead of a generic load.
template <bool SMEM, class T>
NCB_HD T load_prog(const T* p) {
#if defined(__CUDA_ARCH__)
    if (SMEM) {
        static_assert(sizeof(T) == 16 || sizeof(T) == 32 || sizeof(T) == 4 || sizeof(T) == 2, "unsupported size");
        T v;
        const unsigned a = (unsigned)__cvta_generic_to_shared(p);
        if (sizeof(T) == 16) {
            unsigned long long x, y;
            asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(a));
            reinterpret_cast<unsigned long long*>(&v)[0] = x; reinterpret_cast<unsigned long long*>(&v)[1] = y;
        } else if (sizeof(T) == 32) {
            unsigned long long x, y, z, w;
            asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(a));
            asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(z), "=l"(w) : "r"(a + 16));
            unsigned long long* q = reinterpret_cast<unsigned long long*>(&v);
            q[0] = x; q[1] = y; q[2] = z; q[3] = w;
        } else if (sizeof(T) == 4) {
            unsigned x;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(a));
            *reinterpret_cast<unsigned*>(&v) = x;
        } else {
            unsigned short x;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(x) : "r"(a));
            *reinterpret_cast<unsigned short*>(&v) = x;
        }
        return v;
    }
#endif
    return *p;
}

constexpr int MAX_TILE_BITS = 13;   // 2^13 amplitudes * 16 B = 128 KB of shared memory
constexpr int MAX_REG_BITS = 4;     // 2^4 amplitudes per thread
constexpr int MAX_QUBITS = 30;

struct alignas(16) cplx {
    double x, y;
};
NCB_HD cplx cmul(cplx a, cplx b) { return cplx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
NCB_HD cplx cadd(cplx a, cplx b) { return cplx{a.x + b.x, a.y + b.y}; }
NCB_HD cplx cfma(cplx a, cplx b, cplx c) {   // a*b + c
#if defined(__CUDA_ARCH__)
    // four DFMA: written as sums of products the compiler contracts only two of the three operations per component
    // (DMUL + DFMA + DADD), which made the generic 2x2 / 4x4 / 16x16 operators 25 % longer than they have to be
    return cplx{fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y))};
#else
    return cplx{a.x * b.x - a.y * b.y + c.x, a.x * b.y + a.y * b.x + c.y};
#endif
}
NCB_HD double cnorm2(cplx a) { return a.x * a.x + a.y * a.y; }

// ------------------------------------------------------------------------------------------------------------
// Program format
// ------------------------------------------------------------------------------------------------------------
enum OpKind : uint8_t {
    OPK_DIAG = 0,     // diagonal on nb (1..4) tile-local bit positions pos[0..nb); table of 2^nb phases
    OPK_DENSE1 = 1,   // 2x2 on register bit rb[0]
    OPK_DENSE2 = 2,   // 4x4 on register bits rb[0] (matrix bit 0), rb[1] (matrix bit 1)
    OPK_DENSE4 = 3,   // 16x16 on register bits rb[0..3]  (density-matrix superoperators, 3-4 qubit unitaries)
    OPK_RX1 = 4,      // [[a, i b], [i c, d]] with a, b, c, d real: 8 flops per pair.  rx is of this form, sx and x are up
                      // to a global phase (sx = e^{i pi/4} rx(pi/2), x = i rx(pi)), which the compiler factors out.
                      // (Deliberately few kinds: every extra case of the executor's switch costs register shuffling.)
    // Density-matrix superoperators sum_k (K_k U) (x) conj(K_k U) of channels whose Kraus operators map computational basis
    // states to basis states up to a phase / weight (Pauli mixtures, reset, amplitude / phase damping -- every channel of the
    // reference's noise models) composed with a diagonal or permutation-like gate: an element |r><c| only feeds elements
    // with the same r XOR c, so the 16x16 (4x4) matrix is four 4x4 (two 2x2) blocks.  Only in the R = 4 dense kernels.
    OPK_DENSE4BD = 5, // 16x16 on register bits 0..3 (row bits 0,1; column bits 2,3): blocks M[d][r'][r], d = r XOR c; 64 entries
    OPK_DENSE2BD = 6, // 4x4 on register bits rb[0] (row), rb[1] (column): blocks M[d][r'][r]; 8 entries
};

// One executable step inside a pass.  `mat`/`mat_bare` are offsets (in doubles) into the matrix pool: `mat` is the
// gate merged with the channel's base Kraus operator, `mat_bare` the gate alone (used for the last instruction of a
// range that ends in a quantum-jump event, where the operator to apply is only known afterwards).
struct Op {
    uint32_t kind_nb;  // packed kind / nb / tmask / regmask / span / group flag (see the accessors below)
                       // (nb = bits the op acts on; DIAG: tmask = 2^(bits of gather A), regmask = register bits
                       //  the op involves; group headers: see executor)
    uint32_t rbs;      // dense kinds: 8 x 4 bit register-bit indices; DIAG: masks of the tile-index bits of gathers A
                       // (low half) and B (high half)
    int32_t instr;     // index of the instruction this op came from
    int32_t mat;
    int32_t mat_bare;
    uint32_t midx_lo;  // DIAG: multipliers of gathers A and B.  The op's thread bits (<= 6 tile-index bits) are split in two
    uint32_t midx_hi;  //       groups of <= 3; ((jb & mask) * multiplier) >> 29 maps each group's 2^n settings one-to-one onto
                       //       [0, 2^n) (found by search on the host), and the table row is hB * 2^nA + hA (2^nA = tmask field)
    int32_t pad;       // group headers: instruction index of the last member
};
// kind_nb layout: kind [0,4) | nb [4,8) | tmask [8,16) | regmask [16,20) | lean selector [20,25) | span [25,31) |
// group header [31].  Lean selector (R = 3 programs without dense two-qubit operators): the case of the lean executor's
// jump table: 0-7 diagonal with that register mask, 8+b RX1 on register bit b, 11+b DENSE1 on register bit b, and the
// fused rotations the constant-bank fast list is rewritten with (they commute: different bits): 14/15/16 RX1 on bits
// (0,1)/(0,2)/(1,2), 17 RX1 on all three bits; coefficients consecutive in SegConst::coef in ascending bit order.
constexpr int LSEL_X2 = 14, LSEL_X3 = 17;
// rbs: eight 4-bit fields (register-bit indices / tile-local positions are < 16)
NCB_HD int op_kind(const Op& o) { return (int)(o.kind_nb & 0xf); }
NCB_HD int op_nb(const Op& o) { return (int)((o.kind_nb >> 4) & 0xf); }
NCB_HD int op_tmask(const Op& o) { return (int)((o.kind_nb >> 8) & 0xff); }
NCB_HD int op_regmask(const Op& o) { return (int)((o.kind_nb >> 16) & 0xf); }
NCB_HD int op_lean_sel(const Op& o) { return (int)((o.kind_nb >> 20) & 0x1f); }
NCB_HD int op_rb(const Op& o, int b) { return (int)((o.rbs >> (4 * b)) & 0xf); }
// diagonal operators act on up to this many bits (table of 2^bits phases)
NCB_HD constexpr int max_diag_bits(int) { return 6; }

// Bit-field deposit compiled to runs: out = OR_i shift((v & mask[i]), shift[i])  (shift >= 0: left, < 0: right).
constexpr int MAX_RUNS = 14;
struct BitRuns {
    uint32_t mask[MAX_RUNS];
    int8_t shift[MAX_RUNS];
    int8_t n;
    int8_t pad;
};
NCB_HD unsigned apply_runs(const BitRuns& r, unsigned v) {
    unsigned o = 0;
    for (int i = 0; i < r.n; ++i) {
        const unsigned x = v & r.mask[i];
        const int sh = r.shift[i];
        o |= sh >= 0 ? (x << sh) : (x >> (-sh));
    }
    return o;
}

// A pass: every thread pulls 2^R amplitudes out of the tile (the register bits select which), applies the ops whose
// bits are register bits (or diagonal), and writes them back.
struct Pass {
    uint8_t regpos[MAX_REG_BITS];                      // tile-local position of register bit b
    uint8_t tpos[MAX_TILE_BITS];                       // tile-local position of thread-index bit i
    uint8_t pad[3];
    uint16_t regoff[1 << MAX_REG_BITS];                // tile-local index offset of register slot m
    uint16_t regoff_swz[1 << MAX_REG_BITS];            // swz(regoff[m]) (swz is XOR-linear)
    BitRuns truns;                                     // thread index -> tile-local index of its slot 0 (= deposit by tpos)
    int32_t op_begin, op_end;                          // ops of this pass (group headers followed by their members)
    int32_t fop_begin, fop_end;                        // the same work without the members: used when the whole pass runs
    int32_t instr_lo, instr_hi;                        // instruction range covered: [lo, hi)
};

// A segment: a run of consecutive instructions whose qubits all live in one tile (K of the n state bits).
struct Segment {
    uint8_t tile_q[MAX_TILE_BITS];   // state bit of tile-local position p, ascending
    uint8_t pad[3];
    BitRuns eruns;                   // tile-local index (low K-R bits) -> state offset inside the tile
    BitRuns bruns;                   // tile number -> state offset of the tile (deposit into the non-tile state bits)
    uint32_t it_off[1 << MAX_REG_BITS];    // state offset of tile-local index (it << (K-R)), it < 2^R
    uint16_t it_swz[1 << MAX_REG_BITS];    // swz(it << (K-R))
    int32_t pass_begin, pass_end;
    int32_t instr_lo, instr_hi;
};

// Self-contained copy of one segment's program, staged into shared memory by the tile kernel:
//   [SegBlob header][Pass x npass][Op x nops][matrix pool (doubles)], every part 16-byte aligned.
// Pass::op_begin/op_end index the blob's own op array; Op::mat / mat_bare are offsets (doubles) into the blob's pool,
// or, when negative, -(offset + 1) into the program's global pool (16x16 operators are too big to stage).
struct SegBlob {
    Segment seg;
    int32_t npass, nops, pool_doubles, bytes;
    int32_t pass_off, op_off, pool_off, pad;    // byte offsets from the start of the blob
};
constexpr int MAX_BLOB_BYTES = 16 * 1024;

// The hot, warp-uniform part of a segment's program, passed to the tile kernel BY VALUE as a __grid_constant__
// parameter: it then lives in the constant bank and is read with LDC (constant cache) instead of broadcast
// shared-memory loads, which were > 40 % of the LSU wavefronts of the kernel.  Holds the FAST op lists (whole passes,
// no jump event inside) only; partial ranges use the shared-memory blob.
constexpr int SC_MAX_OPS = 64, SC_MAX_PASSES = 12, SC_MAX_COEF = 256;
struct PassHot {
    uint16_t regoff_swz[1 << MAX_REG_BITS];
    int32_t fop_begin, fop_end;        // into SegConst::ops
    int32_t instr_lo, instr_hi;
};
struct SegConst {
    int32_t valid;                     // 0: the segment did not fit; use the blob for everything
    int32_t npass, nops, ncoef;
    // Direct store of the last pass (whole-segment works without events, direct_kernel): the last pass writes its
    // amplitudes from registers to HBM, skipping a shared-memory round trip.  Possible when the L low state bits are the
    // low thread-index bits of that pass (64-byte runs per 4 lanes).  State offset of thread `tid`, register slot m:
    // apply_runs(gruns_last, tid) | goff_last[m].
    int32_t direct;
    // Direct load of the FIRST pass (generated kernels only, ncb_spec.cpp): the same condition and offsets for the first
    // pass -- its amplitudes come from HBM straight into registers, no shared-memory staging of the tile at all.
    int32_t direct_first;
    int32_t pad_[2];
    uint32_t goff_last[1 << MAX_REG_BITS];
    BitRuns gruns_last;
    uint32_t goff_first[1 << MAX_REG_BITS];
    BitRuns gruns_first;
    PassHot pass[SC_MAX_PASSES];
    Op ops[SC_MAX_OPS];                // mat: rx1 / dense1 -> index into coef (4 / 8 doubles); other kinds -> as in the blob
    double coef[SC_MAX_COEF];
};

// Static description of one noise channel (a Kraus list) for the event planner.
struct Channel {
    int32_t dim;          // 2 or 4
    int32_t count;        // number of Kraus operators
    int32_t kraus;        // offset (doubles) of the scaled operators in the matrix pool (count * dim*dim complex)
    int32_t mk;           // offset (doubles) of M_k = K_k^+ K_k, UNSCALED (count * dim*dim complex)
    int32_t bounds;       // offset (doubles) of lo[count-1], hi[count-1]
    int32_t base;         // outcome merged into the gate
    int32_t scale;        // offset (doubles) of the per-operator static scale (count doubles)
    int32_t pad;
};

// Per-instruction static data.
struct Instr {
    int32_t channel;      // -1: no noise step
    int32_t q0, q1;       // state bits the channel acts on (q1 = -1 for one-qubit channels)
    int32_t seg;          // segment that executes it
};

// One unit of work of a tile-kernel launch: run the ops of `seg` whose instruction index is in [instr_lo, instr_hi)
// on state slot `slot`.
enum WorkFlags : int32_t {
    WF_BARE_LAST = 1,      // instruction instr_hi-1 ends in a jump event: apply the bare gate (no base Kraus operator)
    WF_PRO_STATIC = 2,     // first apply Kraus operator ev_k of instruction ev_instr (branch known from u alone)
    WF_PRO_DYNAMIC = 4,    // first apply the operator the decide kernel chose for this slot
    WF_REDUCE = 8,         // finally accumulate the reduced density matrix of the qubits of instruction instr_hi-1
    WF_INIT = 16,          // the state is |0...0>: do not read it
    WF_NORM = 32,          // finally accumulate the squared norm
    WF_MID = 64,           // a jump whose branch is known from u alone happens INSIDE the range: run [instr_lo, mid_instr]
                           // with the bare gate, apply Kraus operator (flags >> 16) of mid_instr, go on with the rest
    WF_PULL = 128,         // (with WF_MID) that in-line jump is the trajectory's ONLY event in this segment: a merged operator
                           // group that does not touch the jump's qubits may run as a whole on one side of the jump (operators
                           // on disjoint qubits commute exactly, the branch does not depend on the state) instead of being
                           // split into its members around it.  Honoured by the generated event kernels, ignored elsewhere.
};
struct Work {
    int32_t slot, src, instr_lo, instr_hi;     // src: slot the tiles are READ from (== slot except when a trajectory
                                               // forks from the shared no-jump prefix, see ncb_schedule.h)
    int32_t flags, ev_instr, ev_k, mid_instr;
};
struct DecideItem {
    int32_t slot, instr;
    double u;
};
struct Decision {
    double scale;
    int32_t k, pad;
};

// ------------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter = (instruction, traj_lo, traj_hi, stream), key = seed
// ------------------------------------------------------------------------------------------------------------
NCB_HD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// Uniform in [0,1) with 53 random bits for (seed, trajectory, instruction).
NCB_HD double philox_uniform(uint64_t seed, uint64_t traj, uint32_t instr) {
    uint32_t c[4] = {instr, (uint32_t)traj, (uint32_t)(traj >> 32), 0x6e636232u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t bits = ((uint64_t)c[0] << 21) ^ ((uint64_t)c[1] >> 11);
    return (double)(bits & ((1ull << 53) - 1)) * (1.0 / 9007199254740992.0);
}

// ------------------------------------------------------------------------------------------------------------
// Event classification: which Kraus branch does uniform u select?
// cp[k](psi) = <C_k>/<C_total> lies in [lo[k], hi[k]] for every state; the reference picks the first k with
// cp[k] >= u (std::discrete_distribution / lower_bound).  Returns kmin (branch >= kmin for sure) and kmax (branch
// <= kmax for sure); kmin == kmax means the branch is known without looking at the state.
// ------------------------------------------------------------------------------------------------------------
constexpr double EVENT_MARGIN = 1e-9;
NCB_HD void classify_uniform(const double* lo, const double* hi, int count, double u, int* kmin, int* kmax) {
    int a = 0;
    while (a < count - 1 && hi[a] < u - EVENT_MARGIN) ++a;          // boundaries certainly below u
    int b = a;
    while (b < count - 1 && !(lo[b] >= u + EVENT_MARGIN)) ++b;       // first boundary certainly >= u
    *kmin = a; *kmax = b;
}

// ------------------------------------------------------------------------------------------------------------
// Tile layout in shared memory: amplitude with tile-local index j lives at slot swz(j).  The XOR fold makes the
// 8 lanes of a quarter-warp hit 8 different 16-byte bank groups for (almost) every choice of register bits.
// ------------------------------------------------------------------------------------------------------------
NCB_HD int swz(int j) { return j ^ (((j >> 3) ^ (j >> 6) ^ (j >> 9) ^ (j >> 12)) & 7); }

// deposit the low bits of v at the positions listed in pos[0..nbits)
NCB_HD int deposit(int v, const uint8_t* pos, int nbits) {
    int r = 0;
    for (int i = 0; i < nbits; ++i) r |= ((v >> i) & 1) << pos[i];
    return r;
}

// ------------------------------------------------------------------------------------------------------------
// Register-blocked pass executor (one thread's share)
//
// Every operator kernel reads the thread's 2^R amplitudes from `in` and writes all of them to `out`.  The kernels
// that carry the 4x4 / 16x16 kinds ping-pong between two register sets (those operators need all inputs for every
// output); the one-bit and diagonal kernels are alias-safe (in == out) and the lean kernels run them in place.
// ------------------------------------------------------------------------------------------------------------
// streaming (evict-first) 128-bit global accesses: every amplitude crosses HBM once per sweep
NCB_HD cplx ld_stream(const cplx* p) {
#if defined(__CUDA_ARCH__)
    cplx v;
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
#else
    return *p;
#endif
}
NCB_HD void st_stream(cplx* p, cplx v) {
#if defined(__CUDA_ARCH__)
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
#else
    *p = v;
#endif
}
template <int R, int B>
NCB_HD void dense1_bit(const cplx* in, cplx* out, const cplx* M) {      // M: four values in registers
    const cplx m00 = M[0], m01 = M[1], m10 = M[2], m11 = M[3];
#pragma unroll
    for (int g = 0; g < (1 << (R - 1)); ++g) {
        const int i0 = ((g >> B) << (B + 1)) | (g & ((1 << B) - 1));
        const int i1 = i0 | (1 << B);
        const cplx x = in[i0], y = in[i1];
        out[i0] = cfma(m01, y, cmul(m00, x));
        out[i1] = cfma(m11, y, cmul(m10, x));
    }
}
template <int R, int B>
NCB_HD void rx1_bit(const cplx* in, cplx* out, const cplx* M) {         // M: four values in registers
    const double ma = M[0].x, mb = M[1].y, mc = M[2].y, md = M[3].x;
#pragma unroll
    for (int g = 0; g < (1 << (R - 1)); ++g) {
        const int i0 = ((g >> B) << (B + 1)) | (g & ((1 << B) - 1));
        const int i1 = i0 | (1 << B);
        const cplx x = in[i0], y = in[i1];
        out[i0] = cplx{ma * x.x - mb * y.y, ma * x.y + mb * y.x};
        out[i1] = cplx{md * y.x - mc * x.y, md * y.y + mc * x.x};
    }
}
// 4x4 on register bits (B0 = matrix bit 0, B1 = matrix bit 1), B0 != B1
template <int R, int B0, int B1>
NCB_HD void dense2_bits(const cplx* in, cplx* out, const cplx* __restrict__ M) {
    constexpr int LO = B0 < B1 ? B0 : B1, HI = B0 < B1 ? B1 : B0;
#pragma unroll
    for (int g = 0; g < (1 << (R - 2)); ++g) {
        int base = ((g >> LO) << (LO + 1)) | (g & ((1 << LO) - 1));
        base = ((base >> HI) << (HI + 1)) | (base & ((1 << HI) - 1));
        const int idx[4] = {base, base | (1 << B0), base | (1 << B1), base | (1 << B0) | (1 << B1)};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            cplx v = cmul(M[4 * r], in[idx[0]]);
            v = cfma(M[4 * r + 1], in[idx[1]], v);
            v = cfma(M[4 * r + 2], in[idx[2]], v);
            v = cfma(M[4 * r + 3], in[idx[3]], v);
            out[idx[r]] = v;
        }
        NCB_NO_HOIST();
    }
}
// 16x16 on all four register bits (the matrix is stored with matrix bit b == register bit b)
template <int R>
NCB_HD void dense4(const cplx* in, cplx* out, const cplx* __restrict__ M) {
    if (R < 4) return;
    constexpr int MASK = (1 << R) - 1;   // keeps the R < 4 instantiations (never executed) in bounds
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        cplx v = cmul(M[16 * r], in[0]);
#pragma unroll
        for (int c = 1; c < 16; ++c) v = cfma(M[16 * r + c], in[c & MASK], v);
        out[r & MASK] = v;
    }
}

// Block-diagonal 16x16: amplitude index m = r | (c << 2) (r: row bits = register bits 0,1; c: column bits = register
// bits 2,3); block d = r ^ c acts on the four amplitudes (r, r ^ d).  A quarter of the arithmetic of dense4.
template <int R>
NCB_HD void dense4bd(const cplx* in, cplx* out, const cplx* __restrict__ M) {
    if (R < 4) return;
    constexpr int MASK = (1 << R) - 1;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
#pragma unroll
        for (int r2 = 0; r2 < 4; ++r2) {
            cplx v = cmul(M[d * 16 + r2 * 4], in[(0 | ((0 ^ d) << 2)) & MASK]);
#pragma unroll
            for (int r = 1; r < 4; ++r) v = cfma(M[d * 16 + r2 * 4 + r], in[(r | ((r ^ d) << 2)) & MASK], v);
            out[(r2 | ((r2 ^ d) << 2)) & MASK] = v;
        }
        NCB_NO_HOIST();
    }
}
// Block-diagonal 4x4 on register bits B0 (row bit) and B1 (column bit): block d = r ^ c on the two amplitudes (r, r ^ d).
template <int R, int B0, int B1>
NCB_HD void dense2bd_bits(const cplx* in, cplx* out, const cplx* __restrict__ M) {
    constexpr int LO = B0 < B1 ? B0 : B1, HI = B0 < B1 ? B1 : B0;
    const cplx m[8] = {M[0], M[1], M[2], M[3], M[4], M[5], M[6], M[7]};
#pragma unroll
    for (int g = 0; g < (1 << (R - 2)); ++g) {
        int base = ((g >> LO) << (LO + 1)) | (g & ((1 << LO) - 1));
        base = ((base >> HI) << (HI + 1)) | (base & ((1 << HI) - 1));
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const int i0 = base | (d << B1), i1 = base | (1 << B0) | ((1 ^ d) << B1);      // (r, c) = (0, d), (1, 1 ^ d)
            const cplx x = in[i0], y = in[i1];
            out[i0] = cfma(m[d * 4 + 1], y, cmul(m[d * 4 + 0], x));
            out[i1] = cfma(m[d * 4 + 3], y, cmul(m[d * 4 + 2], x));
        }
    }
}

// Diagonal operator.  MASK = the register bits it involves: the 2^popc(MASK) phases a thread needs are adjacent in
// the table (the host lays the table out as [thread-bit hash][register bits in MASK]), so one index computation and
// popc(MASK) vector loads serve the 2^R amplitudes.
NCB_HD constexpr int popc_c(int v) { return v ? (v & 1) + popc_c(v >> 1) : 0; }
// bits of m selected by MASK, packed towards bit 0 (pext)
NCB_HD constexpr int pext_c(int m, int MASK) {
    int r = 0, k = 0;
    for (int b = 0; b < 8; ++b)
        if ((MASK >> b) & 1) { r |= ((m >> b) & 1) << k; ++k; }
    return r;
}
// column swizzle key of diagonal-table row h when a row has 2^nr columns (see diag_mask)
NCB_HD constexpr unsigned diag_col_key(unsigned h, int nr) { return nr == 0 ? 0u : (h >> (nr < 3 ? 3 - nr : 0)) & ((1u << nr) - 1u); }
// 128-bit load from a table known to be in shared memory (the generic-address form costs an address-space check per load)
NCB_HD cplx ld_shared_cplx(const cplx* p) {
#if defined(__CUDA_ARCH__)
    cplx v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
#else
    return *p;
#endif
}
template <int R, int MASK, bool SHARED_T = false>
NCB_HD void diag_mask(const Op& op, int jb, const cplx* __restrict__ T, const cplx* in, cplx* out) {
    constexpr int NA = 1 << R, NR = popc_c(MASK);
    // the thread's part of the table index: two multiply-gathers of <= 3 tile bits each (see the Op comment)
    const unsigned x = (unsigned)jb;
    const unsigned hA = ((x & (op.rbs & 0xffffu)) * op.midx_lo) >> 29;
    const unsigned hB = ((x & (op.rbs >> 16)) * op.midx_hi) >> 29;
    const unsigned h = hB * (unsigned)op_tmask(op) + hA;
    const cplx* Tb = T + (h << NR);
    // the 2^NR columns of a row are stored XOR-ed with the row bits that alias rows onto the same shared-memory banks
    // (rows are 2^NR x 16 B, a quarter warp reads 8 x 16 B): lanes on different rows then read different banks
    const unsigned key = diag_col_key(h, NR);
    cplx ph[1 << NR];
#pragma unroll
    for (int r = 0; r < (1 << NR); ++r) ph[r] = SHARED_T ? ld_shared_cplx(Tb + (r ^ key)) : Tb[r ^ key];
#pragma unroll
    for (int m = 0; m < NA; ++m) out[m] = cmul(in[m], ph[pext_c(m, MASK)]);
}

constexpr int FEAT_DENSE2 = 1, FEAT_DENSE4 = 2;   // which operator kinds a kernel instantiation carries

// Applies one op: out = op(in).  One flat switch on (kind, register bits) so that the dispatch is a single jump.
// M: table / matrix in memory (diag, dense2, dense4); m4: the 2x2 of the one-bit kinds, already in registers.
template <int R, int FEAT>
NCB_HD void apply_op(const Op& op, int jb, const cplx* __restrict__ M, const cplx* m4, const cplx* in, cplx* out) {
    constexpr int NA = 1 << R;
    const int kind = op_kind(op);
    const int b0 = op_rb(op, 0), b1 = op_rb(op, 1);
    int sel;
    if (kind == OPK_DIAG) sel = (kind << 4) | op_regmask(op);
    else if (kind == OPK_DENSE4 || kind == OPK_DENSE4BD) sel = kind << 4;
    else if (kind == OPK_DENSE2 || kind == OPK_DENSE2BD) sel = (kind << 4) | (b0 << 2) | b1;
    else sel = (kind << 4) | b0;
    (void)NA;
    switch (sel) {
#define NCB_DIAG(MK) case (OPK_DIAG << 4) | MK: if (MK < (1 << R)) diag_mask<R, (MK < (1 << R) ? MK : 0)>(op, jb, M, in, out); else NCB_UNREACHABLE(); break;
        NCB_DIAG(0) NCB_DIAG(1) NCB_DIAG(2) NCB_DIAG(3) NCB_DIAG(4) NCB_DIAG(5) NCB_DIAG(6) NCB_DIAG(7)
        NCB_DIAG(8) NCB_DIAG(9) NCB_DIAG(10) NCB_DIAG(11) NCB_DIAG(12) NCB_DIAG(13) NCB_DIAG(14) NCB_DIAG(15)
#undef NCB_DIAG
#define NCB_ONE_BIT(B)                                                                   \
    case (OPK_DENSE1 << 4) | B: if (B < R) dense1_bit<R, (B < R ? B : 0)>(in, out, m4); else NCB_UNREACHABLE(); break; \
    case (OPK_RX1 << 4) | B: if (B < R) rx1_bit<R, (B < R ? B : 0)>(in, out, m4); else NCB_UNREACHABLE(); break;
        NCB_ONE_BIT(0) NCB_ONE_BIT(1) NCB_ONE_BIT(2) NCB_ONE_BIT(3)
#undef NCB_ONE_BIT
#define NCB_TWO_BIT(X, Y)                                                                \
    case (OPK_DENSE2 << 4) | (X << 2) | Y:                                               \
        if ((FEAT & FEAT_DENSE2) && R > X && R > Y) dense2_bits<R, (R > X ? X : 0), (R > Y ? Y : 1)>(in, out, M); \
        else NCB_UNREACHABLE();                                                          \
        break;
        NCB_TWO_BIT(0, 1) NCB_TWO_BIT(0, 2) NCB_TWO_BIT(0, 3) NCB_TWO_BIT(1, 0) NCB_TWO_BIT(1, 2) NCB_TWO_BIT(1, 3)
        NCB_TWO_BIT(2, 0) NCB_TWO_BIT(2, 1) NCB_TWO_BIT(2, 3) NCB_TWO_BIT(3, 0) NCB_TWO_BIT(3, 1) NCB_TWO_BIT(3, 2)
#undef NCB_TWO_BIT
    case OPK_DENSE4 << 4: if (FEAT & FEAT_DENSE4) dense4<R>(in, out, M); else NCB_UNREACHABLE(); break;
    case OPK_DENSE4BD << 4: if (FEAT & FEAT_DENSE4) dense4bd<R>(in, out, M); else NCB_UNREACHABLE(); break;
#define NCB_TWO_BIT_BD(X, Y)                                                             \
    case (OPK_DENSE2BD << 4) | (X << 2) | Y:                                             \
        if ((FEAT & FEAT_DENSE4) && R > X && R > Y) dense2bd_bits<R, (R > X ? X : 0), (R > Y ? Y : 1)>(in, out, M); \
        else NCB_UNREACHABLE();                                                          \
        break;
        NCB_TWO_BIT_BD(0, 1) NCB_TWO_BIT_BD(0, 2) NCB_TWO_BIT_BD(0, 3) NCB_TWO_BIT_BD(1, 0) NCB_TWO_BIT_BD(1, 2) NCB_TWO_BIT_BD(1, 3)
        NCB_TWO_BIT_BD(2, 0) NCB_TWO_BIT_BD(2, 1) NCB_TWO_BIT_BD(2, 3) NCB_TWO_BIT_BD(3, 0) NCB_TWO_BIT_BD(3, 1) NCB_TWO_BIT_BD(3, 2)
#undef NCB_TWO_BIT_BD
    default:
        NCB_UNREACHABLE();        // the host compiler only emits the kinds above
#if !defined(__CUDA_ARCH__)
        for (int m = 0; m < NA; ++m) out[m] = in[m];
#endif
        break;
    }
}

// Executes the ops of pass `ps` whose instruction index is in [lo, hi) on this thread's 2^R amplitudes (jb: the thread's
// base tile index in this pass).  bare_last: the op of instruction hi-1 uses the bare gate matrix.
// USE_SC: when the whole pass runs (the common case), read the pass record, the ops and the 2x2 coefficients from the
// constant bank (sc.pass[sc_pass], see SegConst) instead of memory.
// gdst (optional, SegConst::direct): the pass writes its amplitudes to global memory at gdst[gj | sc.goff_last[m]]
// instead of the shared-memory tile; after_load() runs once the pass holds its amplitudes in registers (the direct
// kernel refills the tile buffer with the next tile from there).
struct NoHook { NCB_HD void operator()() const {} };
// ONLY_CST: the caller guarantees the constant-bank fast list applies (whole pass, valid SegConst): the general op
// lists are not compiled in.
template <int R, int FEAT, bool USE_SC, class AfterLoad = NoHook, bool ONLY_CST = false>
NCB_HD void run_pass_thread(cplx* sm, int K, int jb, const Pass& ps, const Op* __restrict__ ops,
                            const double* __restrict__ pool, const double* __restrict__ pool_global, int lo, int hi, bool bare_last,
                            const SegConst& sc, int sc_pass, cplx* __restrict__ gdst = nullptr, unsigned gj = 0,
                            AfterLoad after_load = AfterLoad()) {
    constexpr int NA = 1 << R;
    (void)K;
    const int jbs = swz(jb);
    const bool use_sc = USE_SC && sc.valid != 0;
    const int p_lo = use_sc ? sc.pass[sc_pass].instr_lo : ps.instr_lo, p_hi = use_sc ? sc.pass[sc_pass].instr_hi : ps.instr_hi;
    // the whole pass lies inside [lo, hi) and none of its instructions is the bare one: take the short op list
    const bool fast = lo <= p_lo && hi >= p_hi && !(bare_last && hi == p_hi);
    const bool cst = ONLY_CST ? true : (fast && use_sc);
    cplx a[NA];
    // shared-memory address of register slot m; recomputed at the store (one LOP each) rather than kept in 2^R
    // registers across the op loop -- the laundered copy of jbs stops the compiler from keeping the first set alive
    auto slot_addr = [&](int base, int m) { return base ^ (cst ? sc.pass[sc_pass].regoff_swz[m] : ps.regoff_swz[m]); };
    int jbs_st = jbs;
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+r"(jbs_st));
#endif
#pragma unroll
    for (int m = 0; m < NA; ++m) a[m] = sm[slot_addr(jbs, m)];
    after_load();
    int o = cst ? sc.pass[sc_pass].fop_begin : (fast ? ps.fop_begin : ps.op_begin);
    const int o_end = cst ? sc.pass[sc_pass].fop_end : (fast ? ps.fop_end : ps.op_end);
    Op op;
    const cplx* M = nullptr;
    cplx m4[4];
    // next op to execute at or after position o (range filter, group headers: see ncb_compile.cpp)
    auto fetch = [&]() -> bool {
        for (; o < o_end; ++o) {
            if (cst) {
                op = sc.ops[o];
                const int kind = op_kind(op);
                if (kind == OPK_RX1) {
                    m4[0].x = sc.coef[op.mat]; m4[1].y = sc.coef[op.mat + 1]; m4[2].y = sc.coef[op.mat + 2]; m4[3].x = sc.coef[op.mat + 3];
                } else if (kind == OPK_DENSE1) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { m4[i].x = sc.coef[op.mat + 2 * i]; m4[i].y = sc.coef[op.mat + 2 * i + 1]; }
                } else {
                    M = reinterpret_cast<const cplx*>(op.mat >= 0 ? pool + op.mat : pool_global + (-op.mat - 1));
                }
                ++o;
                return true;
            }
            op = ops[o];
            int moff = op.mat;
            if (!fast) {
                if (op.kind_nb >> 31) {
                    // group header: the product of the next `span` ops; usable when the whole group is inside
                    // [lo, hi) and its last instruction is not the bare one
                    const int span = (int)((op.kind_nb >> 25) & 0x3f), last = op.pad;
                    if (!(op.instr >= lo && (last + 1 < hi || (last + 1 == hi && !bare_last)))) continue;
                    o += span;
                } else {
                    if (op.instr < lo || op.instr >= hi) continue;
                    if (bare_last && op.instr == hi - 1) moff = op.mat_bare;
                }
            }
            M = reinterpret_cast<const cplx*>(moff >= 0 ? pool + moff : pool_global + (-moff - 1));
            const int kind = op_kind(op);
            if (kind == OPK_RX1 || kind == OPK_DENSE1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) m4[i] = M[i];
            }
            ++o;
            return true;
        }
        return false;
    };
    if constexpr (R == 3 && FEAT == 0) {
        if (cst) {
            // hot loop of the lean kernels: the ops come from the constant bank and carry their jump-table case
            for (; o < o_end; ++o) {
                const Op& cop = sc.ops[o];
                const int mat = cop.mat;
                switch (op_lean_sel(cop)) {
#define NCB_LEAN_DIAG(MK) case MK: diag_mask<R, MK, true>(cop, jb, reinterpret_cast<const cplx*>(pool + mat), a, a); break;   /* phase tables are always in the staged pool */
                    NCB_LEAN_DIAG(0) NCB_LEAN_DIAG(1) NCB_LEAN_DIAG(2) NCB_LEAN_DIAG(3)
                    NCB_LEAN_DIAG(4) NCB_LEAN_DIAG(5) NCB_LEAN_DIAG(6) NCB_LEAN_DIAG(7)
#undef NCB_LEAN_DIAG
#define NCB_LEAN_ONE(B)                                                                                              \
    case 8 + B:                                                                                                      \
        m4[0].x = sc.coef[mat]; m4[1].y = sc.coef[mat + 1]; m4[2].y = sc.coef[mat + 2]; m4[3].x = sc.coef[mat + 3];  \
        rx1_bit<R, B>(a, a, m4);                                                                                     \
        break;                                                                                                       \
    case 11 + B:                                                                                                     \
        for (int i = 0; i < 4; ++i) { m4[i].x = sc.coef[mat + 2 * i]; m4[i].y = sc.coef[mat + 2 * i + 1]; }          \
        dense1_bit<R, B>(a, a, m4);                                                                                  \
        break;
                    NCB_LEAN_ONE(0) NCB_LEAN_ONE(1) NCB_LEAN_ONE(2)
#undef NCB_LEAN_ONE
#define NCB_LEAN_RX(B, AT) m4[0].x = sc.coef[mat + AT]; m4[1].y = sc.coef[mat + AT + 1]; m4[2].y = sc.coef[mat + AT + 2]; m4[3].x = sc.coef[mat + AT + 3]; rx1_bit<R, B>(a, a, m4);
                case LSEL_X2 + 0: NCB_LEAN_RX(0, 0) NCB_LEAN_RX(1, 4) break;
                case LSEL_X2 + 1: NCB_LEAN_RX(0, 0) NCB_LEAN_RX(2, 4) break;
                case LSEL_X2 + 2: NCB_LEAN_RX(1, 0) NCB_LEAN_RX(2, 4) break;
                case LSEL_X3: NCB_LEAN_RX(0, 0) NCB_LEAN_RX(1, 4) NCB_LEAN_RX(2, 8) break;
#undef NCB_LEAN_RX
                default: NCB_UNREACHABLE(); break;
                }
            }
            if (gdst) {
#pragma unroll
                for (int m = 0; m < NA; ++m) st_stream(gdst + (gj | sc.goff_last[m]), a[m]);
            } else {
#pragma unroll
                for (int m = 0; m < NA; ++m) sm[slot_addr(jbs_st, m)] = a[m];
            }
            return;
        }
    }
    if constexpr (ONLY_CST && R == 3 && FEAT == 0) {
        return;                                   // unreachable: the lean loop above returned
    } else if constexpr (FEAT == 0) {
        // one-bit and diagonal kernels only: they are alias-safe, so the ops run in place (half the amplitude registers)
        while (fetch()) apply_op<R, FEAT>(op, jb, M, m4, a, a);
    } else {
        cplx b[NA];
        while (true) {
            if (!fetch()) break;
            apply_op<R, FEAT>(op, jb, M, m4, a, b);
            if (!fetch()) {
#pragma unroll
                for (int m = 0; m < NA; ++m) a[m] = b[m];
                break;
            }
            apply_op<R, FEAT>(op, jb, M, m4, b, a);
        }
    }
    if (gdst) {
#pragma unroll
        for (int m = 0; m < NA; ++m) st_stream(gdst + (gj | sc.goff_last[m]), a[m]);
    } else {
#pragma unroll
        for (int m = 0; m < NA; ++m) sm[slot_addr(jbs_st, m)] = a[m];
    }
}

// ------------------------------------------------------------------------------------------------------------
// Simple (non register-blocked) tile loops used by the rare jump events
// ------------------------------------------------------------------------------------------------------------
// Applies a dim x dim (dim = 2 or 4) operator, scaled by `scale`, on tile-local bits p0 (matrix bit 0) and p1.
NCB_HD void tile_apply_dense(cplx* sm, int K, int tid, int nthreads, int dim, int p0, int p1,
                             const cplx* __restrict__ M, double scale) {
    if (dim == 2) {
        for (int g = tid; g < (1 << (K - 1)); g += nthreads) {
            const int i0 = ((g >> p0) << (p0 + 1)) | (g & ((1 << p0) - 1)), i1 = i0 | (1 << p0);
            const cplx x = sm[swz(i0)], y = sm[swz(i1)];
            cplx u = cfma(M[1], y, cmul(M[0], x)), v = cfma(M[3], y, cmul(M[2], x));
            u.x *= scale; u.y *= scale; v.x *= scale; v.y *= scale;
            sm[swz(i0)] = u; sm[swz(i1)] = v;
        }
    } else {
        const int lo = p0 < p1 ? p0 : p1, hi = p0 < p1 ? p1 : p0;
        for (int g = tid; g < (1 << (K - 2)); g += nthreads) {
            int base = ((g >> lo) << (lo + 1)) | (g & ((1 << lo) - 1));
            base = ((base >> hi) << (hi + 1)) | (base & ((1 << hi) - 1));
            const int idx[4] = {base, base | (1 << p0), base | (1 << p1), base | (1 << p0) | (1 << p1)};
            const cplx s0 = sm[swz(idx[0])], s1 = sm[swz(idx[1])], s2 = sm[swz(idx[2])], s3 = sm[swz(idx[3])];
            for (int r = 0; r < 4; ++r) {
                cplx v = cmul(M[4 * r], s0);
                v = cfma(M[4 * r + 1], s1, v);
                v = cfma(M[4 * r + 2], s2, v);
                v = cfma(M[4 * r + 3], s3, v);
                v.x *= scale; v.y *= scale;
                sm[swz(idx[r])] = v;
            }
        }
    }
}

// This thread's contribution to the LOWER TRIANGLE of the reduced density matrix of tile positions p0 (index bit 0)
// and p1 (index bit 1; unused for D = 2): acc[2*(r*D+c)], acc[2*(r*D+c)+1] += psi_r conj(psi_c) for r >= c.
template <int D>
NCB_HD void reduce_rho_thread(const cplx* sm, int K, int tid, int nthreads, int p0, int p1, double* acc) {
    if (D == 2) {
        for (int g = tid; g < (1 << (K - 1)); g += nthreads) {
            const int i0 = ((g >> p0) << (p0 + 1)) | (g & ((1 << p0) - 1)), i1 = i0 | (1 << p0);
            const cplx s[2] = {sm[swz(i0)], sm[swz(i1)]};
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) {
                    acc[2 * (r * 2 + c)] += s[r].x * s[c].x + s[r].y * s[c].y;
                    acc[2 * (r * 2 + c) + 1] += s[r].y * s[c].x - s[r].x * s[c].y;
                }
        }
    } else {
        const int lo = p0 < p1 ? p0 : p1, hi = p0 < p1 ? p1 : p0;
        for (int g = tid; g < (1 << (K - 2)); g += nthreads) {
            int base = ((g >> lo) << (lo + 1)) | (g & ((1 << lo) - 1));
            base = ((base >> hi) << (hi + 1)) | (base & ((1 << hi) - 1));
            const cplx s[4] = {sm[swz(base)], sm[swz(base | (1 << p0))], sm[swz(base | (1 << p1))],
                               sm[swz(base | (1 << p0) | (1 << p1))]};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) {
                    acc[2 * (r * 4 + c)] += s[r].x * s[c].x + s[r].y * s[c].y;
                    acc[2 * (r * 4 + c) + 1] += s[r].y * s[c].x - s[r].x * s[c].y;
                }
        }
    }
}

// rho (lower triangle, dimension d) -> full Hermitian matrix in place
NCB_HD void mirror_rho(double* rho, int d) {
    for (int r = 0; r < d; ++r)
        for (int c = r + 1; c < d; ++c) {
            rho[2 * (r * d + c)] = rho[2 * (c * d + r)];
            rho[2 * (r * d + c) + 1] = -rho[2 * (c * d + r) + 1];
        }
}

// position of state bit q inside the tile
NCB_HD int tile_pos(const Segment& sg, int K, int q) {
    for (int p = 0; p < K; ++p)
        if (sg.tile_q[p] == q) return p;
    return -1;
}

// Branch decision from a reduced density matrix (rho[2*(r*dim+c)] layout as above, any normalisation):
// p_k = Re tr(M_k rho); picks the first k with cumsum(p/sum)[k] >= u (std::discrete_distribution semantics) and
// returns the scale 1/sqrt(p_k) that renormalises K_k psi.  `kraus_scale[k]` is the static scale the stored
// operator already carries.
NCB_HD int decide_branch(const double* rho, const Channel& ch, const double* __restrict__ pool, double u, double* scale_out) {
    const int d = ch.dim, dd = d * d;
    const double* mk = pool + ch.mk;
    double sum = 0.0;
    for (int k = 0; k < ch.count; ++k) {
        double p = 0.0;
        for (int r = 0; r < d; ++r)
            for (int c = 0; c < d; ++c)   // tr(M rho) = sum_{r,c} M[r][c] rho[c][r]
                p += mk[2 * (k * dd + r * d + c)] * rho[2 * (c * d + r)] - mk[2 * (k * dd + r * d + c) + 1] * rho[2 * (c * d + r) + 1];
        sum += p > 0.0 ? p : 0.0;
    }
    int pick = ch.count - 1;
    double cp = 0.0, ppick = 0.0;
    for (int k = 0; k < ch.count; ++k) {
        double p = 0.0;
        for (int r = 0; r < d; ++r)
            for (int c = 0; c < d; ++c)
                p += mk[2 * (k * dd + r * d + c)] * rho[2 * (c * d + r)] - mk[2 * (k * dd + r * d + c) + 1] * rho[2 * (c * d + r) + 1];
        if (p < 0.0) p = 0.0;
        cp += p / sum;
        ppick = p;
        if (k == ch.count - 1 || !(cp < u)) { pick = k; break; }
    }
    // stored operator = static_scale * K_k; we want K_k psi / sqrt(p_k)
    const double st = pool[ch.scale + pick];
    *scale_out = ppick > 0.0 ? 1.0 / (st * sqrt(ppick)) : 0.0;
    return pick;
}

}  // namespace ncb
