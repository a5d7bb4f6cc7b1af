// Building blocks of the SPECIALISED sweep kernels (ncb_spec.cpp generates the kernels of a program from them).
//
// The direct kernel interprets a segment's constant-bank program: per operator a dispatch, constant-bank loads of its
// coefficients and index arithmetic on run-time geometry.  A specialised kernel is the same code with the segment's
// program written out: every pass is a straight line of the hand-written operator kernels of ncb_common.h called with
// LITERAL masks, offsets and bit positions, so the compiler folds the interpreter away (no dispatch, address arithmetic
// with immediates).  Gate coefficients are NOT literals: they arrive in a __grid_constant__ kernel parameter, i.e. in the
// constant bank, where FP64 instructions take them as direct operands -- the generated code, and with it the cubin cache
// key, depends on the structure of the circuit only, so a parameter sweep over one gate skeleton compiles once.
//
//   ncb_seg_<k>      works that run the WHOLE segment k without a jump event: first pass loaded from HBM straight into
//                    registers, last pass stored from registers straight to HBM, shared memory only for the exchanges
//                    between passes.
//   ncb_seg_<k>_ev   works that carry a jump event (pieces [lo, hi) of the segment, in-line jumps, reduced-density
//                    reductions): the pass programs with every operator guarded by its instruction index.
#pragma once
#define NCB_TEMPLATES_ONLY
#include "ncb_kernels.cuh"

namespace ncb {

struct SpecCoef { double c[SC_MAX_COEF]; };     // SegConst::coef of the segment

// diagonal operator with literal gather parameters (see diag_mask); T: the phase table in the staged pool
template <int R, int MASK>
__device__ __forceinline__ void spec_diag(cplx* a, int jb, const cplx* T, unsigned mask_a, unsigned magic_a, unsigned mask_b,
                                          unsigned magic_b, unsigned mult) {
    Op op{};
    op.kind_nb = mult << 8;
    op.rbs = mask_a | (mask_b << 16);
    op.midx_lo = magic_a;
    op.midx_hi = magic_b;
    diag_mask<R, MASK, true>(op, jb, T, a, a);
}
template <int R, int B>
__device__ __forceinline__ void spec_rx(cplx* a, double c0, double c1, double c2, double c3) {
    const cplx m4[4] = {cplx{c0, 0.0}, cplx{0.0, c1}, cplx{0.0, c2}, cplx{c3, 0.0}};
    rx1_bit<R, B>(a, a, m4);
}
// [[0, i b], [i c, 0]]: the x gate (i rx(pi)) merged with a diagonal base operator -- half the arithmetic of spec_rx
template <int R, int B>
__device__ __forceinline__ void spec_xl(cplx* a, double b, double c) {
#pragma unroll
    for (int g = 0; g < (1 << (R - 1)); ++g) {
        const int i0 = ((g >> B) << (B + 1)) | (g & ((1 << B) - 1));
        const int i1 = i0 | (1 << B);
        const cplx x = a[i0], y = a[i1];
        a[i0] = cplx{-b * y.y, b * y.x};
        a[i1] = cplx{-c * x.y, c * x.x};
    }
}
// Rotations with ONE real coefficient factored out (the generator multiplies the factors of a segment into one scalar that
// the last pass applies once): half the FP64 instructions of spec_rx for the symmetric forms.
//   spec_rx_s1: [[1, i t], [i t, 1]]      (a = d, b = c, |a| >= |b|: factor a, t = b / a)
//   spec_rx_s2: [[t, i], [i, t]]          (a = d, b = c, |a| <  |b|: factor b, t = a / b)
//   spec_rx_a:  [[1, i tb], [i tc, td]]   (general: factor a)
//   spec_xs:    [[0, i], [i, 0]]          (x-like with b = c: no arithmetic at all)
//   spec_xl2:   [[0, i], [i t, 0]]        (x-like, factor b, t = c / b)
template <int R, int B>
__device__ __forceinline__ void spec_rx_s1(cplx* a, double t) {
#pragma unroll
    for (int g = 0; g < (1 << (R - 1)); ++g) {
        const int i0 = ((g >> B) << (B + 1)) | (g & ((1 << B) - 1)), i1 = i0 | (1 << B);
        const cplx x = a[i0], y = a[i1];
        a[i0] = cplx{x.x - t * y.y, x.y + t * y.x};
        a[i1] = cplx{y.x - t * x.y, y.y + t * x.x};
    }
}
template <int R, int B>
__device__ __forceinline__ void spec_rx_s2(cplx* a, double t) {
#pragma unroll
    for (int g = 0; g < (1 << (R - 1)); ++g) {
        const int i0 = ((g >> B) << (B + 1)) | (g & ((1 << B) - 1)), i1 = i0 | (1 << B);
        const cplx x = a[i0], y = a[i1];
        a[i0] = cplx{t * x.x - y.y, t * x.y + y.x};
        a[i1] = cplx{t * y.x - x.y, t * y.y + x.x};
    }
}
template <int R, int B>
__device__ __forceinline__ void spec_rx_a(cplx* a, double tb, double tc, double td) {
#pragma unroll
    for (int g = 0; g < (1 << (R - 1)); ++g) {
        const int i0 = ((g >> B) << (B + 1)) | (g & ((1 << B) - 1)), i1 = i0 | (1 << B);
        const cplx x = a[i0], y = a[i1];
        a[i0] = cplx{x.x - tb * y.y, x.y + tb * y.x};
        a[i1] = cplx{td * y.x - tc * x.y, td * y.y + tc * x.x};
    }
}
template <int R, int B>
__device__ __forceinline__ void spec_xs(cplx* a) {
#pragma unroll
    for (int g = 0; g < (1 << (R - 1)); ++g) {
        const int i0 = ((g >> B) << (B + 1)) | (g & ((1 << B) - 1)), i1 = i0 | (1 << B);
        const cplx x = a[i0], y = a[i1];
        a[i0] = cplx{-y.y, y.x};
        a[i1] = cplx{-x.y, x.x};
    }
}
template <int R, int B>
__device__ __forceinline__ void spec_xl2(cplx* a, double t) {
#pragma unroll
    for (int g = 0; g < (1 << (R - 1)); ++g) {
        const int i0 = ((g >> B) << (B + 1)) | (g & ((1 << B) - 1)), i1 = i0 | (1 << B);
        const cplx x = a[i0], y = a[i1];
        a[i0] = cplx{-y.y, y.x};
        a[i1] = cplx{-t * x.y, t * x.x};
    }
}
template <int R, int B>
__device__ __forceinline__ void spec_u(cplx* a, double c0, double c1, double c2, double c3, double c4, double c5, double c6, double c7) {
    const cplx m4[4] = {cplx{c0, c1}, cplx{c2, c3}, cplx{c4, c5}, cplx{c6, c7}};
    dense1_bit<R, B>(a, a, m4);
}
// generic 2x2 with its first entry factored out: [[1, p], [q, r]] (p, q, r complex): 12 instead of 16 FP64 instructions per pair
template <int R, int B>
__device__ __forceinline__ void spec_u1(cplx* a, double pr, double pi, double qr, double qi, double rr, double ri) {
    const cplx p{pr, pi}, q{qr, qi}, r{rr, ri};
#pragma unroll
    for (int g = 0; g < (1 << (R - 1)); ++g) {
        const int i0 = ((g >> B) << (B + 1)) | (g & ((1 << B) - 1)), i1 = i0 | (1 << B);
        const cplx x = a[i0], y = a[i1];
        a[i0] = cfma(p, y, x);
        a[i1] = cfma(r, y, cmul(q, x));
    }
}
// the same with the 2x2 (row-major, interleaved re/im) read from the staged pool in shared memory (event kernels)
template <int R, int B>
__device__ __forceinline__ void spec_rx_m(cplx* a, const double* m) {
    const cplx* M = reinterpret_cast<const cplx*>(m);
    const cplx m4[4] = {ld_shared_cplx(M), ld_shared_cplx(M + 1), ld_shared_cplx(M + 2), ld_shared_cplx(M + 3)};
    rx1_bit<R, B>(a, a, m4);
}
template <int R, int B>
__device__ __forceinline__ void spec_u_m(cplx* a, const double* m) {
    const cplx* M = reinterpret_cast<const cplx*>(m);
    const cplx m4[4] = {ld_shared_cplx(M), ld_shared_cplx(M + 1), ld_shared_cplx(M + 2), ld_shared_cplx(M + 3)};
    dense1_bit<R, B>(a, a, m4);
}

}  // namespace ncb
