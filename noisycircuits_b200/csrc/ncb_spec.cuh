// Building blocks of the SPECIALISED sweep kernels (ncb_spec.cpp generates one kernel per segment from them).
//
// The direct kernel interprets a segment's constant-bank program: per operator a dispatch, constant-bank loads of its
// coefficients and index arithmetic on run-time geometry.  A specialised kernel is the same code with the segment's
// program written out: every pass is a straight line of the hand-written operator kernels of ncb_common.h called with
// LITERAL coefficients, masks and offsets, so the compiler folds the interpreter away (no dispatch, no LDCU, address
// arithmetic with immediates).  Same tile geometry, same shared-memory layout, same phase tables as the direct kernel:
// results are bit-identical to it.
#pragma once
#define NCB_TEMPLATES_ONLY
#include "ncb_kernels.cuh"

namespace ncb {

// diagonal operator with literal gather parameters (see diag_mask); T: the phase table in the staged pool
template <int MASK>
__device__ __forceinline__ void spec_diag(cplx* a, int jb, const cplx* T, unsigned mask_a, unsigned magic_a, unsigned mask_b,
                                          unsigned magic_b, unsigned mult) {
    Op op{};
    op.kind_nb = mult << 8;
    op.rbs = mask_a | (mask_b << 16);
    op.midx_lo = magic_a;
    op.midx_hi = magic_b;
    diag_mask<3, MASK, true>(op, jb, T, a, a);
}
template <int B>
__device__ __forceinline__ void spec_rx(cplx* a, double c0, double c1, double c2, double c3) {
    const cplx m4[4] = {cplx{c0, 0.0}, cplx{0.0, c1}, cplx{0.0, c2}, cplx{c3, 0.0}};
    rx1_bit<3, B>(a, a, m4);
}
template <int B>
__device__ __forceinline__ void spec_u(cplx* a, double c0, double c1, double c2, double c3, double c4, double c5, double c6, double c7) {
    const cplx m4[4] = {cplx{c0, c1}, cplx{c2, c3}, cplx{c4, c5}, cplx{c6, c7}};
    dense1_bit<3, B>(a, a, m4);
}

}  // namespace ncb
