#include "ncb_spec.h"

#include <cstddef>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <sstream>

namespace ncb {

namespace {

std::string lit(double v) {            // exact literal
    char b[64];
    std::snprintf(b, sizeof b, "%a", v);
    return b;
}
std::string hex(uint32_t v) {
    char b[32];
    std::snprintf(b, sizeof b, "0x%xu", v);
    return b;
}
// apply_runs(r, v) written out
std::string runs_expr(const BitRuns& r, const std::string& v) {
    if (r.n == 0) return "0u";
    std::ostringstream o;
    o << "(";
    for (int i = 0; i < r.n; ++i) {
        if (i) o << " | ";
        const int sh = r.shift[i];
        o << "((" << v << " & " << hex(r.mask[i]) << ")";
        if (sh > 0) o << " << " << sh;
        if (sh < 0) o << " >> " << -sh;
        o << ")";
    }
    o << ")";
    return o.str();
}

struct BlobView {
    const SegBlob* blob;
    const Segment* sg;
    const Pass* passes;
    const Op* ops;
};
BlobView view_blob(const Program& P, int seg) {
    const unsigned char* braw = P.blobs.data() + P.blob_off[seg];
    BlobView v;
    v.blob = reinterpret_cast<const SegBlob*>(braw);
    v.sg = &v.blob->seg;
    v.passes = reinterpret_cast<const Pass*>(braw + v.blob->pass_off);
    v.ops = reinterpret_cast<const Op*>(braw + v.blob->op_off);
    return v;
}

bool lean_tiled(const Program& P) {
    return P.mode != MODE_DENSITY && (P.R == 3 || P.R == 4) && !P.has_dense2 && !P.has_dense4 && P.nbits > P.K;
}

// the fast list of pass p (whole pass, no jump inside): coefficients as literals or C.c[..], tables in the staged pool
// scalar: the product of the real factors taken out of the rotations so far (literal mode, factor_scalars); the caller
// multiplies it back in once, in the segment's last pass
bool emit_fast_ops(std::ostringstream& o, const SegConst& sc, int p, int R, bool literal, double* scalar = nullptr /* re, im */) {
    auto scale_by = [&](double re, double im) {
        const double x = scalar[0] * re - scalar[1] * im, y = scalar[0] * im + scalar[1] * re;
        scalar[0] = x; scalar[1] = y;
    };
    const PassHot& ph = sc.pass[p];
    auto coef = [&](int i) { return literal ? lit(sc.coef[i]) : "C.c[" + std::to_string(i) + "]"; };
    for (int i = ph.fop_begin; i < ph.fop_end; ++i) {
        const Op& op = sc.ops[i];
        const int kind = op_kind(op), ls = op_lean_sel(op);
        auto rx = [&](int bit, int at) {
            // literal coefficients: a rotation with a vanishing diagonal (x = i rx(pi): cos(pi/2) = 6e-17) is written as the
            // anti-diagonal operator it is -- half the FP64 instructions; the dropped terms are below 1e-15 of the kept ones
            const double* c = sc.coef + at;          // [[c0, i c1], [i c2, c3]]
            const bool xlike = std::fabs(c[0]) <= 1e-15 * std::fabs(c[1]) && std::fabs(c[3]) <= 1e-15 * std::fabs(c[2]);
            if (literal && scalar && std::isfinite(c[0] + c[1] + c[2] + c[3])) {
                // one real coefficient factored out: half the FP64 instructions for the symmetric forms, none for x
                const std::string hd = "<" + std::to_string(R) + ", " + std::to_string(bit) + ">(a";
                const bool sym = std::fabs(c[0] - c[3]) <= 1e-15 * std::fabs(c[0]) && std::fabs(c[1] - c[2]) <= 1e-15 * std::fabs(c[1]);
                if (xlike && c[1] != 0.0) {
                    scale_by(c[1], 0.0);
                    if (std::fabs(c[1] - c[2]) <= 1e-15 * std::fabs(c[1])) o << "            spec_xs" << hd << ");\n";
                    else o << "            spec_xl2" << hd << ", " << lit(c[2] / c[1]) << ");\n";
                    return;
                }
                if (sym && (c[0] != 0.0 || c[1] != 0.0)) {
                    if (std::fabs(c[0]) >= std::fabs(c[1])) { scale_by(c[0], 0.0); o << "            spec_rx_s1" << hd << ", " << lit(c[1] / c[0]) << ");\n"; }
                    else { scale_by(c[1], 0.0); o << "            spec_rx_s2" << hd << ", " << lit(c[0] / c[1]) << ");\n"; }
                    return;
                }
                if (std::fabs(c[0]) >= 0.25 * std::max(std::fabs(c[1]), std::max(std::fabs(c[2]), std::fabs(c[3])))) {
                    scale_by(c[0], 0.0);
                    o << "            spec_rx_a" << hd << ", " << lit(c[1] / c[0]) << ", " << lit(c[2] / c[0]) << ", " << lit(c[3] / c[0]) << ");\n";
                    return;
                }
            }
            if (literal && xlike) {
                o << "            spec_xl<" << R << ", " << bit << ">(a, " << lit(c[1]) << ", " << lit(c[2]) << ");\n";
                return;
            }
            o << "            spec_rx<" << R << ", " << bit << ">(a, " << coef(at) << ", " << coef(at + 1) << ", " << coef(at + 2) << ", " << coef(at + 3) << ");\n";
        };
        if (kind == OPK_DIAG) {
            if (op.mat < 0 || (op.mat & 1)) return false;      // (phase tables are complex-aligned entries of the staged pool)
            o << "            spec_diag<" << R << ", " << op_regmask(op) << ">(a, jb" << p << ", pool + " << op.mat / 2 << ", " << hex(op.rbs & 0xffffu) << ", "
              << hex(op.midx_lo) << ", " << hex(op.rbs >> 16) << ", " << hex(op.midx_hi) << ", " << (unsigned)op_tmask(op) << "u);\n";
        } else if (kind == OPK_RX1 && R == 3 && ls == LSEL_X3) {      // fused rotations of the lean fast list (R = 3 only)
            rx(0, op.mat); rx(1, op.mat + 4); rx(2, op.mat + 8);
        } else if (kind == OPK_RX1 && R == 3 && ls >= LSEL_X2 && ls < LSEL_X3) {
            const int pair = ls - LSEL_X2;                      // (0,1) (0,2) (1,2)
            rx(pair == 2 ? 1 : 0, op.mat); rx(pair == 0 ? 1 : 2, op.mat + 4);
        } else if (kind == OPK_RX1) {
            rx(op_rb(op, 0), op.mat);
        } else if (kind == OPK_DENSE1) {
            const double* c = sc.coef + op.mat;      // u00, u01, u10, u11 (re, im)
            const double m0 = std::hypot(c[0], c[1]), mx = std::max(std::hypot(c[2], c[3]), std::max(std::hypot(c[4], c[5]), std::hypot(c[6], c[7])));
            if (literal && scalar && std::isfinite(m0 + mx) && m0 >= 0.25 * mx && m0 > 0.0) {
                // first entry factored out (a complex scalar this time): [[1, p], [q, r]]
                auto div = [&](int i, double* re, double* im) {
                    const double d = c[0] * c[0] + c[1] * c[1];
                    *re = (c[i] * c[0] + c[i + 1] * c[1]) / d; *im = (c[i + 1] * c[0] - c[i] * c[1]) / d;
                };
                double pr, pi, qr, qi, rr, ri;
                div(2, &pr, &pi); div(4, &qr, &qi); div(6, &rr, &ri);
                scale_by(c[0], c[1]);
                o << "            spec_u1<" << R << ", " << op_rb(op, 0) << ">(a, " << lit(pr) << ", " << lit(pi) << ", " << lit(qr) << ", " << lit(qi) << ", "
                  << lit(rr) << ", " << lit(ri) << ");\n";
                continue;
            }
            o << "            spec_u<" << R << ", " << op_rb(op, 0) << ">(a";
            for (int k = 0; k < 8; ++k) o << ", " << coef(op.mat + k);
            o << ");\n";
        } else {
            return false;
        }
    }
    return true;
}

// ---- whole-segment kernel -----------------------------------------------------------------------------------------
bool emit_segment_kernel(std::ostringstream& o, const Program& P, int seg, const SpecOptions& opt) {
    const int K = P.K, R = P.R, NA = 1 << R;
    const SegConst& sc = P.segconst[seg];
    const BlobView bv = view_blob(P, seg);
    const Segment& sg = *bv.sg;
    const int nbuf = opt.nbuf == 1 ? 1 : 2;
    const int load = sc.direct_first ? std::min(std::max(opt.load, 0), 2) : 0;
    const int npass = sc.npass;
    // the last segment of a program whose operators do not preserve the norm also accumulates the squared norm of every
    // tile (WF_NORM; what the event-capable kernel does for works that carry events): the amplitudes are in registers then
    const bool with_norm = seg == (int)P.segs.size() - 1 && !P.norm_preserving;
    double scalar[2] = {1.0, 0.0};
    double* const sp = (opt.literal && opt.factor) ? scalar : nullptr;
    auto emit_norm = [&](std::ostringstream& os) {
        if (scalar[1] != 0.0) {   // the factors taken out of this segment's one-qubit operators, multiplied back in once
            os << "            {\n                const cplx S{" << lit(scalar[0]) << ", " << lit(scalar[1]) << "};\n";
            for (int m = 0; m < NA; ++m) os << "                a[" << m << "] = cmul(a[" << m << "], S);\n";
            os << "            }\n";
        } else if (scalar[0] != 1.0) {
            os << "            {\n                const double S = " << lit(scalar[0]) << ";\n";
            for (int m = 0; m < NA; ++m) os << "                a[" << m << "].x *= S; a[" << m << "].y *= S;\n";
            os << "            }\n";
        }
        if (!with_norm) return;
        os << "            {   // squared norm of the tile: fixed shuffle tree per warp, warps added in order (deterministic)\n"
              "                double nrm = 0.0;\n";
        for (int m = 0; m < NA; ++m) os << "                nrm += cnorm2(a[" << m << "]);\n";
        os << "                nrm = warp_sum(nrm);\n"
              "                if ((tid & 31) == 0) s_nrm[tid >> 5] = nrm;\n"
              "                __syncthreads();\n"
              "                if (tid == 0) {\n"
              "                    double t2 = 0.0;\n"
              "                    for (int wv = 0; wv < (nthreads >> 5); ++wv) t2 += s_nrm[wv];\n"
              "                    norm_part[(long long)works[t >> tile_shift].slot * ntiles + tl] = t2;\n"
              "                }\n"
              "                __syncthreads();\n"
              "            }\n";
    };
    o << "extern \"C\" __global__ void __launch_bounds__(" << (1 << (K - R)) << ", " << spec_ctas_per_sm(P, opt) << ") ncb_seg_" << seg
      << "(const __grid_constant__ SpecCoef C, const unsigned char* __restrict__ pool_g, int pool_bytes, int pool_cap,\n"
         "        const Work* __restrict__ works, int nworks, cplx* __restrict__ states, int tile_shift, double* __restrict__ norm_part) {\n"
         "    extern __shared__ __align__(16) unsigned char smem_raw[];\n"
         "    const int tid = threadIdx.x, nthreads = blockDim.x;\n"
         "    const int ntiles = 1 << tile_shift;\n"
         "    const long long total = (long long)nworks << tile_shift;\n"
         "    cplx* const buf0 = reinterpret_cast<cplx*>(smem_raw + pool_cap);\n"
      << (with_norm ? "    __shared__ double s_nrm[32];\n" : "    (void)norm_part;\n")
      << "    (void)nthreads;\n"
         "    for (int i = tid * 16; i < pool_bytes; i += nthreads * 16) cp_async16(smem_raw + i, pool_g + i);\n"
         "    cp_async_wait_all();\n    __syncthreads();\n"
         "    const cplx* const pool = reinterpret_cast<const cplx*>(smem_raw);\n"
         "    const unsigned utid = (unsigned)tid;\n"
      << "    const unsigned gj_last = " << runs_expr(sc.gruns_last, "utid") << ";\n";
    if (load == 0) {
        // ---- tile staged in shared memory with cp.async (as in the interpreting direct kernel) ----
        o << "    const unsigned offA = " << runs_expr(sg.eruns, "utid") << ";\n"
          << "    const int sA = swz(tid);\n"
          << "    auto issue_load = [&](long long tt, cplx* sm) {\n"
             "        const unsigned tl = (unsigned)(tt & (ntiles - 1));\n"
          << "        const cplx* st = states + ((long long)works[tt >> tile_shift].src << " << P.nbits << ") + " << runs_expr(sg.bruns, "tl") << ";\n";
        for (int it = 0; it < NA; ++it)
            o << "        cp_async16(sm + (sA ^ " << sg.it_swz[it] << "), st + (offA | " << hex(sg.it_off[it]) << "));\n";
        o << "        asm volatile(\"cp.async.commit_group;\\n\" ::: \"memory\");\n    };\n"
             "    long long t = blockIdx.x;\n    if (t >= total) return;\n    issue_load(t, buf0);\n";
        if (nbuf == 2) {
            o << "    for (int iter = 0; t < total; t += gridDim.x, ++iter) {\n"
                 "        cplx* const sm = buf0 + ((size_t)(iter & 1) << " << K << ");\n"
                 "        const long long tn = t + gridDim.x;\n"
                 "        if (tn < total) issue_load(tn, buf0 + ((size_t)((iter + 1) & 1) << " << K << "));\n"
                 "        else asm volatile(\"cp.async.commit_group;\\n\" ::: \"memory\");\n"
                 "        asm volatile(\"cp.async.wait_group 1;\\n\" ::: \"memory\");\n";
        } else {
            o << "    for (; t < total; t += gridDim.x) {\n"
                 "        cplx* const sm = buf0;\n"
                 "        const long long tn = t + gridDim.x;\n"
                 "        asm volatile(\"cp.async.wait_group 0;\\n\" ::: \"memory\");\n";
        }
        o << "        const unsigned tl = (unsigned)(t & (ntiles - 1));\n"
          << "        cplx* const dst = states + ((long long)works[t >> tile_shift].slot << " << P.nbits << ") + " << runs_expr(sg.bruns, "tl") << ";\n"
          << "        __syncthreads();\n";
        for (int p = 0; p < npass; ++p) {
            const PassHot& ph = sc.pass[p];
            const bool last = p == npass - 1;
            o << "        {   // pass " << p << "\n            cplx a[" << NA << "];\n"
              << "            const int jb" << p << " = (int)" << runs_expr(bv.passes[p].truns, "utid") << ", jbs" << p << " = swz(jb" << p << ");\n";
            for (int m = 0; m < NA; ++m) o << "            a[" << m << "] = sm[jbs" << p << " ^ " << ph.regoff_swz[m] << "];\n";
            if (last && nbuf == 1) o << "            __syncthreads();\n            if (tn < total) issue_load(tn, buf0);\n";
            if (last && nbuf == 2) o << "            __syncthreads();            // this buffer is refilled at the top of the next iteration\n";
            if (!emit_fast_ops(o, sc, p, R, opt.literal != 0, sp)) return false;
            if (last) {
                emit_norm(o);
                for (int m = 0; m < NA; ++m) o << "            st_stream(dst + (gj_last | " << hex(sc.goff_last[m]) << "), a[" << m << "]);\n";
            } else {
                for (int m = 0; m < NA; ++m) o << "            sm[jbs" << p << " ^ " << ph.regoff_swz[m] << "] = a[" << m << "];\n";
                o << "            __syncthreads();\n";
            }
            o << "        }\n";
        }
        o << "    }\n}\n\n";
        return true;
    }
    // ---- first pass loaded from HBM straight into registers ----
    // Shared memory only carries the exchanges between passes (a thread reads back the slots it wrote within a pass, so one
    // barrier per pass TRANSITION).  nbuf = 2: tiles alternate between two buffers, so the first store of a tile never meets
    // the last reads of the previous one; nbuf = 1: one more barrier per tile instead.
    o << "    const unsigned gj_first = " << runs_expr(sc.gruns_first, "utid") << ";\n"
      << "    auto gload = [&](long long tt, cplx* r) {\n"
         "        const unsigned tl = (unsigned)(tt & (ntiles - 1));\n"
      << "        const cplx* st = states + ((long long)works[tt >> tile_shift].src << " << P.nbits << ") + (" << runs_expr(sg.bruns, "tl") << " | gj_first);\n";
    for (int m = 0; m < NA; ++m) o << "        r[" << m << "] = ld_stream(st + " << hex(sc.goff_first[m]) << ");\n";
    o << "    };\n"
         "    long long t = blockIdx.x;\n    if (t >= total) return;\n"
      << "    cplx a[" << NA << "];\n    gload(t, a);\n"
      << "    for (int iter = 0; t < total; t += gridDim.x, ++iter) {\n";
    if (npass > 1) {
        if (nbuf == 2) o << "        cplx* const sm = buf0 + ((size_t)(iter & 1) << " << K << ");\n";
        else o << "        cplx* const sm = buf0;\n";
    }
    o << "        const long long tn = t + gridDim.x;\n"
         "        const unsigned tl = (unsigned)(t & (ntiles - 1));\n"
      << "        cplx* const dst = states + ((long long)works[t >> tile_shift].slot << " << P.nbits << ") + " << runs_expr(sg.bruns, "tl") << ";\n";
    if (load == 2) o << "        cplx nx[" << NA << "];\n";
    for (int p = 0; p < npass; ++p) {
        const PassHot& ph = sc.pass[p];
        const bool last = p == npass - 1;
        o << "        {   // pass " << p << "\n"
          << "            const int jb" << p << " = (int)" << runs_expr(bv.passes[p].truns, "utid") << ", jbs" << p << " = swz(jb" << p << ");\n"
          << "            (void)jbs" << p << ";\n";
        if (p > 0)
            for (int m = 0; m < NA; ++m) o << "            a[" << m << "] = sm[jbs" << p << " ^ " << ph.regoff_swz[m] << "];\n";
        if (last) {
            if (nbuf == 1 && npass > 1) o << "            __syncthreads();            // every thread holds its amplitudes: the buffer is free for the next tile\n";
            if (load == 2) o << "            if (tn < total) gload(tn, nx);\n";
        }
        if (!emit_fast_ops(o, sc, p, R, opt.literal != 0, sp)) return false;
        if (last) {
            emit_norm(o);
            for (int m = 0; m < NA; ++m) o << "            st_stream(dst + (gj_last | " << hex(sc.goff_last[m]) << "), a[" << m << "]);\n";
        } else {
            for (int m = 0; m < NA; ++m) o << "            sm[jbs" << p << " ^ " << ph.regoff_swz[m] << "] = a[" << m << "];\n";
            o << "            __syncthreads();\n";
        }
        o << "        }\n";
    }
    if (load == 2) {
        o << "        if (tn < total) {\n";
        for (int m = 0; m < NA; ++m) o << "            a[" << m << "] = nx[" << m << "];\n";
        o << "        }\n";
    } else {
        o << "        if (tn < total) gload(tn, a);\n";
    }
    o << "    }\n}\n\n";
    return true;
}

// ---- event kernel ---------------------------------------------------------------------------------------------------
// One op of the general list (the blob's op array), guarded / selected as the interpreter's fetch() does.
void emit_general_op(std::ostringstream& o, const Op& op, int p, int R, const std::string& off_expr, const char* ind) {
    const int kind = op_kind(op);
    if (kind == OPK_DIAG) {
        o << ind << "spec_diag<" << R << ", " << op_regmask(op) << ">(a, jb" << p << ", pool + ((" << off_expr << ") >> 1), " << hex(op.rbs & 0xffffu) << ", "
          << hex(op.midx_lo) << ", " << hex(op.rbs >> 16) << ", " << hex(op.midx_hi) << ", " << (unsigned)op_tmask(op) << "u);\n";
    } else if (kind == OPK_RX1) {
        o << ind << "spec_rx_m<" << R << ", " << op_rb(op, 0) << ">(a, pool_d + (" << off_expr << "));\n";
    } else {
        o << ind << "spec_u_m<" << R << ", " << op_rb(op, 0) << ">(a, pool_d + (" << off_expr << "));\n";
    }
}

// The pass programs of a segment with every operator guarded by its instruction index (what run_range + fetch() interpret):
// the body of a function / lambda with (lo, hi, bare, qb, qe) in scope, jb<p> / jbs<p>, pool and pool_d declared by the caller.
// qb / qe: tile positions of the qubits of the in-line jump this range begins after / ends in when merged groups that do
// not touch them may stay whole (WF_PULL); ~0u otherwise (every group "touches": strict splitting by instruction index).
// ckpt_pass0 >= 0 (resident trunk): before pass p the tile is copied to ckpt + ((ckpt_pass0 + p) << K) when ckpt != nullptr.
// sc / fopt (event kernels of tiled programs): a pass that lies entirely inside the range and holds no bare operator runs its
// FAST list -- the text of the whole-segment kernel (merged groups, literal / factored coefficients, generic products split) --
// with the factors taken out of its rotations multiplied back at the end of the pass; only the pass a jump cuts is guarded.
void emit_guarded_passes(std::ostringstream& o, const BlobView& bv, int R, int K = 0, int ckpt_pass0 = -1, const SegConst* sc = nullptr,
                         const SpecOptions* fopt = nullptr) {
    const int NA = 1 << R, npass = bv.blob->npass;
    for (int p = 0; p < npass; ++p) {
        const Pass& ps = bv.passes[p];
        if (ps.instr_hi <= ps.instr_lo) continue;
        o << "        if (" << ps.instr_lo << " < hi && " << ps.instr_hi << " > lo) {   // pass " << p << "\n";
        if (ckpt_pass0 >= 0)
            o << "            if (ckpt != nullptr) {\n                cplx* const c = ckpt + ((size_t)" << ckpt_pass0 + p << " << " << K << ");\n"
              << "                for (int i = (int)threadIdx.x; i < " << (1 << K) << "; i += (int)blockDim.x) c[i] = sm[i];\n"
              << "                __syncthreads();      // (the copy reads every thread's slots: nobody may store this pass' results before it is done)\n            }\n";
        bool with_fast = false;
        if (sc && fopt && sc->valid && p < sc->npass) {
            std::ostringstream f;
            double scalar[2] = {1.0, 0.0};
            if (emit_fast_ops(f, *sc, p, R, fopt->literal != 0, (fopt->literal && fopt->factor) ? scalar : nullptr)) {
                with_fast = true;
                o << "            if (lo <= " << ps.instr_lo << " && hi >= " << ps.instr_hi << " && !(bare && hi == " << ps.instr_hi << ")) {\n"
                  << "            cplx a[" << NA << "];\n";
                for (int m = 0; m < NA; ++m) o << "            a[" << m << "] = sm[jbs" << p << " ^ " << ps.regoff_swz[m] << "];\n";
                o << f.str();
                if (scalar[1] != 0.0) {
                    o << "            {\n                const cplx S{" << lit(scalar[0]) << ", " << lit(scalar[1]) << "};\n";
                    for (int m = 0; m < NA; ++m) o << "                a[" << m << "] = cmul(a[" << m << "], S);\n";
                    o << "            }\n";
                } else if (scalar[0] != 1.0) {
                    o << "            {\n                const double S = " << lit(scalar[0]) << ";\n";
                    for (int m = 0; m < NA; ++m) o << "                a[" << m << "].x *= S; a[" << m << "].y *= S;\n";
                    o << "            }\n";
                }
                for (int m = 0; m < NA; ++m) o << "            sm[jbs" << p << " ^ " << ps.regoff_swz[m] << "] = a[" << m << "];\n";
                o << "            __syncthreads();\n            } else {\n";
            }
        }
        o << "            cplx a[" << NA << "];\n";
        for (int m = 0; m < NA; ++m) o << "            a[" << m << "] = sm[jbs" << p << " ^ " << ps.regoff_swz[m] << "];\n";
        // (hot: the operator stands alone in the list; cold: a member of a merged group, run one by one only when a jump
        //  cuts the group -- the hints keep the cold code out of the fall-through path and of the instruction cache)
        auto member = [&](const Op& op, const char* ind, bool hot) {
            o << ind << "if (" << (hot ? "__builtin_expect(" : "(") << op.instr << " >= lo && " << op.instr << " < hi" << (hot ? ", 1)" : ")") << ") {\n";
            std::string off = std::to_string(op.mat);
            if (op.mat_bare != op.mat) off = "(bare && hi == " + std::to_string(op.instr + 1) + ") ? " + std::to_string(op.mat_bare) + " : " + std::to_string(op.mat);
            emit_general_op(o, op, p, R, off, (std::string(ind) + "    ").c_str());
            o << ind << "}\n";
        };
        for (int i = ps.op_begin; i < ps.op_end; ++i) {
            const Op& op = bv.ops[i];
            if (op.kind_nb >> 31) {
                const int span = (int)((op.kind_nb >> 25) & 0x3f), last = op.pad;
                // tile positions the merged operator touches
                unsigned gm = 0;
                if (op_kind(op) == OPK_DIAG) {
                    gm = (op.rbs & 0xffffu) | (op.rbs >> 16);
                    for (int r = 0; r < R; ++r) if ((op_regmask(op) >> r) & 1) gm |= 1u << ps.regpos[r];
                } else {
                    gm = 1u << ps.regpos[op_rb(op, 0)];
                }
                o << "            if (__builtin_expect(" << op.instr << " >= lo && ((" << last + 1 << " < hi || (" << last + 1 << " == hi && !bare)) || ("
                  << op.instr << " < hi && !(" << hex(gm) << " & qe))), 1)) {\n";
                emit_general_op(o, op, p, R, std::to_string(op.mat), "                ");
                o << "            } else if (" << op.instr << " < lo && !(" << hex(gm) << " & qb)) {\n"
                     "                // ran as a whole before the jump this range begins after\n"
                     "            } else {\n";
                for (int j = 1; j <= span; ++j) member(bv.ops[i + j], "                ", false);
                o << "            }\n";
                i += span;
            } else {
                member(op, "            ", true);
            }
        }
        for (int m = 0; m < NA; ++m) o << "            sm[jbs" << p << " ^ " << ps.regoff_swz[m] << "] = a[" << m << "];\n";
        o << "            __syncthreads();\n";
        if (with_fast) o << "            }\n";
        o << "        }\n";
    }
}

bool emit_event_kernel(std::ostringstream& o, const Program& P, int seg, const SpecOptions& opt) {
    const int K = P.K, R = P.R, NA = 1 << R;
    const BlobView bv = view_blob(P, seg);
    const Segment& sg = *bv.sg;
    const int npass = bv.blob->npass;
    for (int i = 0; i < bv.blob->nops; ++i) {
        const Op& op = bv.ops[i];
        const int kind = op_kind(op);
        if (kind != OPK_DIAG && kind != OPK_RX1 && kind != OPK_DENSE1) return false;
        if (op.mat < 0 || op.mat_bare < 0 || (op.mat & 1) || (op.mat_bare & 1)) return false;
    }
    o << "extern \"C\" __global__ void __launch_bounds__(" << (1 << (K - R)) << ", " << spec_ctas_per_sm(P, opt) << ") ncb_seg_" << seg << "_ev"
      << "(const __grid_constant__ SpecCoef C, DevProgram P, const unsigned char* __restrict__ pool_g, int pool_bytes, int pool_cap,\n"
         "        const Work* __restrict__ works, int nworks, cplx* __restrict__ states, int tile_shift,\n"
         "        double* __restrict__ red_part, double* __restrict__ norm_part, Decision* decisions,\n"
         "        const double* __restrict__ work_u, int* tickets, int pull_on) {\n"
         "    extern __shared__ __align__(16) unsigned char smem_raw[];\n"
         "    __shared__ double s_buf[32 * 20];\n"
         "    __shared__ double s_out[20];\n"
         "    __shared__ int s_last;\n"
         "    const int tid = threadIdx.x, nthreads = blockDim.x;\n"
         "    const int ntiles = 1 << tile_shift;\n"
         "    const long long total = (long long)nworks << tile_shift;\n"
         "    cplx* const sm = reinterpret_cast<cplx*>(smem_raw + pool_cap);\n"
         "    for (int i = tid * 16; i < pool_bytes; i += nthreads * 16) cp_async16(smem_raw + i, pool_g + i);\n"
         "    cp_async_wait_all();\n    __syncthreads();\n"
         "    const cplx* const pool = reinterpret_cast<const cplx*>(smem_raw);\n"
         "    const double* const pool_d = reinterpret_cast<const double*>(smem_raw);\n"
         "    (void)C;\n"
      << "    const Segment* const sg = P.segs + " << seg << ";\n"
      << "    const int K = " << K << ";\n"
         "    const unsigned utid = (unsigned)tid;\n"
      << "    const unsigned offA = " << runs_expr(sg.eruns, "utid") << ";\n"
      << "    const int sA = swz(tid);\n";
    for (int p = 0; p < npass; ++p)
        o << "    const int jb" << p << " = (int)" << runs_expr(bv.passes[p].truns, "utid") << ", jbs" << p << " = swz(jb" << p << ");\n";
    o << "    auto issue_load = [&](long long tt, const Work& w) {\n"
         "        if (w.flags & WF_INIT) return;\n"
         "        const unsigned tl = (unsigned)(tt & (ntiles - 1));\n"
      << "        const cplx* st = states + ((long long)w.src << " << P.nbits << ") + " << runs_expr(sg.bruns, "tl") << ";\n";
    for (int it = 0; it < NA; ++it)
        o << "        cp_async16(sm + (sA ^ " << sg.it_swz[it] << "), st + (offA | " << hex(sg.it_off[it]) << "));\n";
    o << "    };\n";
    o << "    auto run_passes = [&](const int lo, const int hi, const bool bare, const unsigned qb, const unsigned qe) {\n";
    emit_guarded_passes(o, bv, R, 0, -1, opt.ev_fast ? &P.segconst[seg] : nullptr, &opt);
    o << "    };\n";
    o << "    long long t = blockIdx.x;\n    if (t >= total) return;\n"
         "    Work w = works[t >> tile_shift];\n"
         "    issue_load(t, w);\n"
         "    asm volatile(\"cp.async.commit_group;\\n\" ::: \"memory\");\n"
         "    for (; t < total; t += gridDim.x) {\n"
         "        const long long tn = t + gridDim.x;\n"
         "        asm volatile(\"cp.async.wait_group 0;\\n\" ::: \"memory\");\n"
         "        __syncthreads();\n"
         "        const int o = (int)(t & (ntiles - 1));\n"
         "        if (w.flags & WF_INIT) {\n";
    for (int it = 0; it < NA; ++it) o << "            sm[sA ^ " << sg.it_swz[it] << "] = cplx{0.0, 0.0};\n";
    o << "            if (o == 0 && tid == 0) sm[swz(0)] = cplx{1.0, 0.0};\n"
         "            __syncthreads();\n"
         "        }\n"
         "        if (w.flags & WF_PRO_STATIC) event_apply(sm, P, sg, w.ev_instr, w.ev_k, 1.0);\n"
         "        else if (w.flags & WF_PRO_DYNAMIC) event_apply(sm, P, sg, w.ev_instr, decisions[w.slot].k, decisions[w.slot].scale);\n"
         "        {\n"
         "            int lo = w.instr_lo;\n"
         "            unsigned qmid = ~0u;\n"
         "            if ((w.flags & WF_PULL) && pull_on) {\n"
         "                const Instr in = P.instrs[w.mid_instr];\n"
         "                qmid = 1u << tile_pos(*sg, K, in.q0);\n"
         "                if (in.q1 >= 0) qmid |= 1u << tile_pos(*sg, K, in.q1);\n"
         "            }\n"
         "            for (int part = (w.flags & WF_MID) ? 0 : 1; part < 2; ++part) {\n"
         "                const int hi = part == 0 ? w.mid_instr + 1 : w.instr_hi;\n"
         "                const bool bare = part == 0 ? true : (w.flags & WF_BARE_LAST) != 0;\n"
         "                const bool after_mid = part == 1 && (w.flags & WF_MID);\n"
         "                run_passes(lo, hi, bare, after_mid ? qmid : ~0u, part == 0 ? qmid : ~0u);\n"
         "                if (part == 0) {\n"
         "                    event_apply(sm, P, sg, w.mid_instr, (int)((unsigned)w.flags >> 16), 1.0);\n"
         "                    lo = w.mid_instr + 1;\n"
         "                }\n"
         "            }\n"
         "        }\n"
         "        if (w.flags & WF_REDUCE) {\n"
         "            const Instr in = P.instrs[w.instr_hi - 1];\n"
         "            const int p0 = tile_pos(*sg, K, in.q0), p1 = in.q1 >= 0 ? tile_pos(*sg, K, in.q1) : -1;\n"
         "            event_reduce(sm, K, p0, p1, s_buf, s_out, red_part + ((long long)w.slot * ntiles + o) * RED_VALUES);\n"
         "            // the CTA that delivers the last partial of this state decides the branch: no decide launch in the chain\n"
         "            // (thread 0 wrote the partial: its fence publishes it before the ticket, the second fence orders the last\n"
         "            //  CTA's reads after it; a fence by every thread of every reducing tile cost 2-3x on these launches)\n"
         "            if (tickets != nullptr) {\n"
         "            if (tid == 0) {\n"
         "                __threadfence();\n"
         "                const int last = atomicAdd(&tickets[w.slot], 1) == ntiles - 1;\n"
         "                __threadfence();\n"
         "                s_last = last;\n"
         "            }\n"
         "            __syncthreads();\n"
         "            if (s_last) {\n"
         "                fused_decide(P, red_part + (long long)w.slot * ntiles * RED_VALUES, ntiles, w.instr_hi - 1, work_u[t >> tile_shift], decisions + w.slot, s_buf);\n"
         "                if (tid == 0) tickets[w.slot] = 0;\n"
         "            }\n"
         "            }\n"
         "        }\n"
         "        if (w.flags & WF_NORM) {\n"
         "            double v[1] = {0.0};\n"
         "            for (int i = tid; i < (1 << K); i += nthreads) v[0] += cnorm2(sm[i]);\n"
         "            block_reduce<1>(v, s_buf, s_out);\n"
         "            if (tid == 0) norm_part[(long long)w.slot * ntiles + o] = s_out[0];\n"
         "        }\n"
         "        {\n"
      << "            cplx* st = states + ((long long)w.slot << " << P.nbits << ") + " << runs_expr(sg.bruns, "(unsigned)o") << ";\n";
    for (int it = 0; it < NA; ++it)
        o << "            st_stream(st + (offA | " << hex(sg.it_off[it]) << "), sm[sA ^ " << sg.it_swz[it] << "]);\n";
    o << "        }\n"
         "        __syncthreads();\n"
         "        if (tn < total) {\n"
         "            w = works[tn >> tile_shift];\n"
         "            issue_load(tn, w);\n"
         "        }\n"
         "        asm volatile(\"cp.async.commit_group;\\n\" ::: \"memory\");\n"
         "    }\n"
         "    asm volatile(\"cp.async.wait_all;\\n\" ::: \"memory\");\n"
         "}\n\n";
    return true;
}

// ---- resident kernels -----------------------------------------------------------------------------------------------
// States of <= 2^K amplitudes (one tile, identity index map): the whole trajectory lives in one CTA's shared memory.  The
// module holds the segments' guarded pass programs as device functions (the same text the event kernels are made of,
// coefficients and phase tables read from the staged pools: structure-only source) and two kernels around them:
//   ncb_res_trunk   one CTA: the jump-free evolution from |0...0>, its state left in global memory at every pass boundary
//   ncb_res         persistent CTAs over the trajectories that leave the trunk: start from the last checkpoint before the
//                   first flagged draw, events handled in line (reduced density matrix, branch decision, Kraus operator)
// which: 0 both kernels (cache key, source query), 1 ncb_res only, 2 ncb_res_trunk only (two translation units compile in parallel;
// the trunk takes the fast path of every segment, so its unit leaves the guarded programs out)
bool emit_resident_module(std::ostringstream& o, const Program& P, const SpecOptions& opt, int which = 0) {
    const int K = P.K, R = P.R, nseg = (int)P.segs.size(), NT = 1 << (K - R);
    const ResidentLayout lay = spec_resident_layout(P, opt);
    for (int seg = 0; seg < nseg; ++seg) {
        const BlobView bv = view_blob(P, seg);
        for (int i = 0; i < bv.blob->nops; ++i) {          // (general lists and fast lists)
            const Op& op = bv.ops[i];
            const int kind = op_kind(op);
            if (kind != OPK_DIAG && kind != OPK_RX1 && kind != OPK_DENSE1) return false;
            if (op.mat < 0 || op.mat_bare < 0 || (op.mat & 1) || (op.mat_bare & 1)) return false;
        }
        o << "static __device__ __noinline__ void res_seg_" << seg << "(cplx* __restrict__ sm, const unsigned char* __restrict__ poolb, const int lo, const int hi,\n"
             "        const bool bare, cplx* __restrict__ ckpt) {\n"
             "    const cplx* const pool = reinterpret_cast<const cplx*>(poolb);\n"
             "    const double* const pool_d = reinterpret_cast<const double*>(poolb);\n"
             "    const unsigned utid = threadIdx.x;\n"
             "    const unsigned qb = ~0u, qe = ~0u;\n"
             "    (void)pool; (void)pool_d; (void)qb; (void)qe; (void)ckpt;\n";
        for (int p = 0; p < bv.blob->npass; ++p)
            o << "    const int jb" << p << " = (int)" << runs_expr(bv.passes[p].truns, "utid") << ", jbs" << p << " = swz(jb" << p << ");\n"
              << "    (void)jb" << p << "; (void)jbs" << p << ";\n";
        if (opt.resident_fast) {
            // the whole segment without a bare operator (no jump event inside, the trunk): the passes' FAST lists -- merged groups
            // only, generic products with their phases handed to the tables -- without any guard
            o << "    if (lo <= " << P.segs[seg].instr_lo << " && hi >= " << P.segs[seg].instr_hi << " && !(bare && hi == " << P.segs[seg].instr_hi << ")) {\n";
            for (int p = 0; p < bv.blob->npass; ++p) {
                const Pass& ps = bv.passes[p];
                if (ps.instr_hi <= ps.instr_lo) continue;
                o << "        {   // pass " << p << "\n"
                  << "            if (ckpt != nullptr) {\n                cplx* const c = ckpt + ((size_t)" << P.segs[seg].pass_begin + p << " << " << K << ");\n"
                  << "                for (int i = (int)threadIdx.x; i < " << (1 << K) << "; i += (int)blockDim.x) c[i] = sm[i];\n"
                  << "                __syncthreads();\n            }\n"
                  << "            cplx a[" << (1 << R) << "];\n";
                for (int m = 0; m < (1 << R); ++m) o << "            a[" << m << "] = sm[jbs" << p << " ^ " << ps.regoff_swz[m] << "];\n";
                for (int i = ps.fop_begin; i < ps.fop_end; ++i) emit_general_op(o, bv.ops[i], p, R, std::to_string(bv.ops[i].mat), "            ");
                for (int m = 0; m < (1 << R); ++m) o << "            sm[jbs" << p << " ^ " << ps.regoff_swz[m] << "] = a[" << m << "];\n";
                o << "            __syncthreads();\n        }\n";
            }
            o << "        return;\n    }\n";
        }
        if (!(which == 2 && opt.resident_fast)) emit_guarded_passes(o, bv, R, K, P.segs[seg].pass_begin);
        o << "}\n\n";
    }
    // [lo, hi) across the segments
    o << "static __device__ __forceinline__ void res_run(cplx* sm, const unsigned char* poolb, const int lo, const int hi, const bool bare, cplx* ckpt) {\n";
    for (int seg = 0; seg < nseg; ++seg)
        o << "    if (" << P.segs[seg].instr_lo << " < hi && " << P.segs[seg].instr_hi << " > lo) res_seg_" << seg << "(sm, poolb + " << lay.pool_off[seg]
          << ", lo, hi, bare, ckpt);\n";
    o << "}\n\n";
    const int nwords = (P.G + 31) / 32;
    o << "namespace {\n"
         "constexpr int RK = " << K << ", RNB = " << P.nbits << ", RDIM = 1 << RNB, RNT = " << NT << ", RG = " << P.G << ", RNPASS = " << (int)P.passes.size()
      << ", RNWORDS = " << nwords << ", RPOOL = " << lay.pool_bytes << ";\n"
         "__device__ __forceinline__ void res_stage(unsigned char* poolb, const ResArgs& A) {\n"
         "    const int tid = threadIdx.x;\n";
    for (int seg = 0; seg < nseg; ++seg)
        o << "    for (int i = tid * 16; i < " << lay.pool_len[seg] << "; i += RNT * 16) cp_async16(poolb + " << lay.pool_off[seg] << " + i, A.blobs + A.seg_pool[" << 2 * seg
          << "] + i);\n";
    o << "    cp_async_wait_all();\n"
         "}\n"
         "__device__ __forceinline__ double res_inv_norm(const cplx* sm, double* s_buf, double* s_out) {\n"
         "    double v[1] = {0.0};\n"
         "    for (int j = threadIdx.x; j < (1 << RK); j += RNT) v[0] += cnorm2(sm[j]);\n"
         "    block_reduce<1>(v, s_buf, s_out);\n"
         "    const double inv = 1.0 / s_out[0];\n"
         "    __syncthreads();\n"
         "    return inv;\n"
         "}\n"
         "}  // namespace\n\n";
    if (which != 1)
    o << "extern \"C\" __global__ void __launch_bounds__(" << NT << ", 1) ncb_res_trunk(DevProgram P, ResArgs A) {\n"
         "    extern __shared__ __align__(16) unsigned char smem_raw[];\n"
         "    __shared__ double s_buf[32 * 20];\n"
         "    __shared__ double s_out[RED_VALUES];\n"
         "    cplx* const sm = reinterpret_cast<cplx*>(smem_raw);\n"
         "    unsigned char* const poolb = smem_raw + ((size_t)16 << RK);\n"
         "    const int tid = threadIdx.x;\n"
         "    (void)P;\n"
         "    res_stage(poolb, A);\n"
         "    for (int j = tid; j < (1 << RK); j += RNT) sm[j] = cplx{0.0, 0.0};\n"
         "    __syncthreads();\n"
         "    if (tid == 0) sm[swz(0)] = cplx{1.0, 0.0};\n"
         "    __syncthreads();\n"
         "    res_run(sm, poolb, 0, RG, false, A.ckpt);\n"
         "    {\n"
         "        cplx* const c = A.ckpt + ((size_t)RNPASS << RK);\n"
         "        for (int i = tid; i < (1 << RK); i += RNT) c[i] = sm[i];\n"
         "    }\n"
         "    const double inv = A.normalize ? res_inv_norm(sm, s_buf, s_out) : 1.0;\n"
         "    for (int j = tid; j < RDIM; j += RNT) A.trunk_probs[j] = cnorm2(sm[swz(j)]) * inv;\n"
         "}\n\n";
    if (which == 2) return true;
    o << "extern \"C\" __global__ void __launch_bounds__(" << NT << ", " << lay.ctas << ") ncb_res(DevProgram P, ResArgs A) {\n"
         "    extern __shared__ __align__(16) unsigned char smem_raw[];\n"
         "    __shared__ double s_buf[32 * 20];\n"
         "    __shared__ double s_out[RED_VALUES];\n"
         "    __shared__ double s_red[20];\n"
         "    __shared__ int s_k;\n"
         "    __shared__ double s_scale;\n"
         "    cplx* const sm = reinterpret_cast<cplx*>(smem_raw);\n"
         "    unsigned char* const poolb = smem_raw + ((size_t)16 << RK);\n"
         "    unsigned* const evmask = reinterpret_cast<unsigned*>(poolb + RPOOL);\n"
         "    const int tid = threadIdx.x;\n"
         "    double* const acc = A.acc_global ? A.partial + (long long)blockIdx.x * RDIM : reinterpret_cast<double*>(evmask + ((RNWORDS + 3) & ~3));\n"
         "    res_stage(poolb, A);\n"
         "    for (int j = tid; j < RDIM; j += RNT) acc[j] = 0.0;\n"
         "    __syncthreads();\n"
         "    const int n_items = A.n_items[0];\n"
         "    if (blockIdx.x == 0 && tid == 0 && A.counters) A.counters[2] = (unsigned long long)(A.count - n_items);\n"
         "    unsigned long long n_events = 0, n_hard = 0;\n"
         "    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {\n"
         "        struct { long long tl; int first; } item = {A.item_traj[it], (int)A.item_first[it]};\n"
         "        const long long tl = item.tl;\n"
         "        const unsigned long long traj = (unsigned long long)(A.traj_begin + tl);\n"
         "        const double* utab = A.uniforms ? A.uniforms + tl * RG : nullptr;\n"
         "        const ResCkpt ck = A.ckpt_for[item.first];\n"
         "        // the trunk's state at the last pass boundary before the first flagged draw\n"
         "        if (ck.pass >= 0) {\n"
         "            const cplx* src = A.ckpt + ((size_t)ck.pass << RK);\n"
         "            for (int i = tid; i < (1 << RK); i += RNT) cp_async16(sm + i, src + i);\n"
         "        } else {\n"
         "            for (int j = tid; j < (1 << RK); j += RNT) sm[j] = cplx{0.0, 0.0};\n"
         "        }\n"
         "        for (int i = tid; i < RNWORDS; i += RNT) evmask[i] = 0u;\n"
         "        __syncthreads();\n"
         "        if (ck.pass < 0 && tid == 0) sm[swz(0)] = cplx{1.0, 0.0};\n"
         "        // which of the later draws leave their base branch?\n"
         "        for (int a = tid; a < P.n_noisy; a += RNT) {\n"
         "            const int i = P.noisy[a];\n"
         "            if (i < item.first) continue;\n"
         "            const double u = utab ? utab[i] : philox_uniform(A.seed, traj, (uint32_t)P.orig[i]);\n"
         "            if (!(u > A.thresh[2 * a] && u <= A.thresh[2 * a + 1])) atomicOr(&evmask[i >> 5], 1u << (i & 31));\n"
         "        }\n"
         "        cp_async_wait_all();\n"
         "        __syncthreads();\n"
         "        int cur = ck.instr;\n"
         "        while (true) {\n"
         "            int e = RG;\n"
         "            for (int wd = cur >> 5; wd < RNWORDS; ++wd) {\n"
         "                unsigned m = evmask[wd];\n"
         "                if (wd == (cur >> 5)) m &= ~0u << (cur & 31);\n"
         "                if (m) { e = (wd << 5) + __ffs(m) - 1; break; }\n"
         "            }\n"
         "            const bool ev = e < RG;\n"
         "            res_run(sm, poolb, cur, ev ? e + 1 : RG, ev, nullptr);\n"
         "            if (!ev) break;\n"
         "            const Instr in = P.instrs[e];\n"
         "            const Channel ch = P.chans[in.channel];\n"
         "            const double u = utab ? utab[e] : philox_uniform(A.seed, traj, (uint32_t)P.orig[e]);\n"
         "            int kmin, kmax;\n"
         "            {   // the channel's bounds through shared memory: one parallel fetch instead of a chain of dependent global loads\n"
         "                const int nb = 2 * (ch.count - 1);\n"
         "                const bool staged = nb <= 32 * 20;\n"
         "                if (staged) {\n"
         "                    for (int i = tid; i < nb; i += RNT) s_buf[i] = P.pool[ch.bounds + i];\n"
         "                    __syncthreads();\n"
         "                }\n"
         "                const double* bnd = staged ? s_buf : P.pool + ch.bounds;\n"
         "                classify_uniform(bnd, bnd + (ch.count - 1), ch.count, u, &kmin, &kmax);\n"
         "                __syncthreads();\n"
         "            }\n"
         "            int k = kmin;\n"
         "            double scale = 1.0;\n"
         "            bool base_after_all = false;\n"
         "            if (kmin != kmax) {\n"
         "                ++n_hard;\n"
         "                event_reduce(sm, RK, in.q0, ch.dim == 4 ? in.q1 : -1, s_buf, s_red, s_out);\n"
         "                if (tid == 0) mirror_rho(s_out, ch.dim);\n"
         "                __syncthreads();\n"
         "                decide_branch_block(s_out, ch, P.pool, u, s_buf, 32 * 20, &s_k, &s_scale);\n"
         "                k = s_k; scale = s_scale;\n"
         "            } else if (k == ch.base) {\n"
         "                base_after_all = true;     // flagged by the cheap test only (u within the margin): base branch\n"
         "            }\n"
         "            if (!base_after_all) ++n_events;\n"
         "            const cplx* M = reinterpret_cast<const cplx*>(P.pool + ch.kraus) + k * ch.dim * ch.dim;\n"
         "            tile_apply_dense(sm, RK, tid, RNT, ch.dim, in.q0, in.q1, M, scale);\n"
         "            __syncthreads();\n"
         "            cur = e + 1;\n"
         "        }\n"
         "        const double inv = A.normalize ? res_inv_norm(sm, s_buf, s_out) : 1.0;\n"
         "        for (int j = tid; j < RDIM; j += RNT) acc[j] += cnorm2(sm[swz(j)]) * inv;\n"
         "        if (A.states_out) {\n"
         "            const double f = sqrt(inv);\n"
         "            for (int j = tid; j < RDIM; j += RNT) {\n"
         "                cplx a = sm[swz(j)];\n"
         "                a.x *= f; a.y *= f;\n"
         "                A.states_out[tl * RDIM + j] = a;\n"
         "            }\n"
         "        }\n"
         "        __syncthreads();\n"
         "    }\n"
         "    if (!A.acc_global)\n"
         "        for (int j = tid; j < RDIM; j += RNT) A.partial[(long long)blockIdx.x * RDIM + j] = acc[j];\n"
         "    if (tid == 0 && A.counters) {\n"
         "        atomicAdd(&A.counters[0], n_events);\n"
         "        atomicAdd(&A.counters[1], n_hard);\n"
         "    }\n"
         "}\n";
    return true;
}

}  // namespace

bool spec_supported(const Program& P, int seg) {
    if (!lean_tiled(P)) return false;
    if (seg < 0 || seg >= (int)P.segconst.size()) return false;
    const SegConst& sc = P.segconst[seg];
    return sc.valid && sc.direct && sc.npass >= 1;
}

bool spec_ev_supported(const Program& P, int seg) {
    if (!lean_tiled(P) || seg < 0 || seg >= (int)P.blob_off.size()) return false;
    // every state offset of the tile has to exist (tiles of a state smaller than a tile carry virtual bits)
    return P.nbits >= P.K && P.segs[seg].pad[0] == 0;          // (not for a segment that is one 5..8-bit unitary)
}

int spec_ctas_per_sm(const Program& P, const SpecOptions& opt) {
    // registers: 64 K per SM; shared memory: nbuf tiles of 16 B x 2^K plus the phase tables (<= 16 KB) within 227 KB
    const int threads = 1 << (P.K - P.R);
    int pool_cap = 0;                                  // what the engine stages per CTA beside the tiles (1 KB granules)
    for (size_t sgi = 0; sgi < P.blob_off.size(); ++sgi) {
        const SegBlob& blob = *reinterpret_cast<const SegBlob*>(P.blobs.data() + P.blob_off[sgi]);
        pool_cap = std::max(pool_cap, (blob.bytes - blob.pool_off + 1023) & ~1023);
    }
    const int nbuf = opt.nbuf == 1 ? 1 : 2;
    if (opt.ctas > 0) return opt.ctas;
    const int by_smem = (int)((227 * 1024) / ((size_t)nbuf * (16u << P.K) + (size_t)pool_cap + 2048));
    const int by_regs = 65536 / (threads * (P.R == 4 ? 128 : 64));
    return std::max(1, std::min({by_smem, by_regs, 8}));
}

void spec_pool(const Program& P, int seg, int* offset, int* bytes) {
    const SegBlob& blob = *reinterpret_cast<const SegBlob*>(P.blobs.data() + P.blob_off[seg]);
    *offset = blob.pool_off;
    *bytes = blob.bytes - blob.pool_off;
}

std::string spec_source(const Program& P, int seg_begin, int seg_end, const SpecOptions& opt) {
    std::ostringstream o;
    o << "// generated by ncb_spec.cpp -- do not edit\n#include \"ncb_spec.cuh\"\nusing namespace ncb;\n";
    for (int seg = seg_begin; seg < seg_end; ++seg) {
        if (spec_supported(P, seg)) {
            std::ostringstream k;
            if (emit_segment_kernel(k, P, seg, opt)) o << k.str();
        }
        if (opt.events && spec_ev_supported(P, seg)) {
            std::ostringstream k;
            if (emit_event_kernel(k, P, seg, opt)) o << k.str();
        }
    }
    return o.str();
}

bool spec_resident_supported(const Program& P) {
    if (P.mode != MODE_MCWF || (P.R != 3 && P.R != 4) || P.has_dense2 || P.has_dense4) return false;
    if (P.nbits > P.K || !P.generics.empty() || P.segs.empty()) return false;
    if (P.blob_off.size() != P.segs.size()) return false;
    for (const Segment& sg : P.segs)
        for (int p = 0; p < P.K; ++p)
            if (sg.tile_q[p] != p) return false;                   // (the resident kernels address the state bits directly)
    return spec_resident_layout(P, SpecOptions()).ctas >= 1;
}

ResidentLayout spec_resident_layout(const Program& P, const SpecOptions& opt) {
    ResidentLayout lay;
    int at = 0;
    for (size_t sgi = 0; sgi < P.blob_off.size(); ++sgi) {
        const SegBlob& blob = *reinterpret_cast<const SegBlob*>(P.blobs.data() + P.blob_off[sgi]);
        lay.pool_src.push_back((int)P.blob_off[sgi] + blob.pool_off);
        lay.pool_len.push_back((blob.bytes - blob.pool_off + 15) & ~15);
        lay.pool_off.push_back(at);
        at += lay.pool_len.back();
    }
    lay.pool_bytes = at;
    const int threads = 1 << (P.K - P.R);
    const size_t fixed = ((size_t)16 << P.K) + (size_t)at + (size_t)((((P.G + 31) / 32) + 3) & ~3) * 4;
    const size_t acc = (size_t)8 << P.nbits, budget = 227 * 1024, stat = 6 * 1024 + 1024;   // static arrays + per-CTA reservation
    const int by_regs = std::max(1, 65536 / (threads * (P.R == 4 ? 128 : 64)));
    auto fit = [&](size_t per) { return (int)std::min<size_t>({budget / (per + stat), (size_t)by_regs, (size_t)8}); };
    // the probability accumulator stays in shared memory unless moving it to global memory (L2) buys another CTA per SM
    const int with_acc = fit(fixed + acc), without = fit(fixed);
    lay.acc_global = without > with_acc ? 1 : 0;
    lay.ctas = lay.acc_global ? without : with_acc;
    if (opt.ctas > 0 && opt.ctas <= lay.ctas) lay.ctas = opt.ctas;
    lay.smem = fixed + (lay.acc_global ? 0 : acc);
    return lay;
}

std::vector<int32_t> spec_resident_ckpts(const Program& P) {
    // checkpoint before (global) pass p: valid when every instruction of the passes before p precedes every instruction of
    // the passes from p on; then it is the state after instructions [0, L_p).  Table: execution position e -> (pass, L) of the
    // last valid checkpoint with L <= e; e = G -> the final state (pass = number of passes).
    const int np = (int)P.passes.size();
    std::vector<int> lo_min(np + 1, P.G), hi_max(np + 1, 0);
    std::vector<char> nonempty(np, 0);
    for (size_t sgi = 0; sgi < P.segs.size(); ++sgi) {
        const BlobView bv = view_blob(P, (int)sgi);
        for (int p = 0; p < bv.blob->npass; ++p) {
            const Pass& ps = bv.passes[p];
            const int gp = P.segs[sgi].pass_begin + p;
            if (gp >= np) continue;
            int mn = P.G, mx = -1;
            for (int i = ps.op_begin; i < ps.op_end; ++i) {
                const Op& op = bv.ops[i];
                mn = std::min(mn, (int)op.instr);
                mx = std::max(mx, (op.kind_nb >> 31) ? (int)op.pad : (int)op.instr);
            }
            nonempty[gp] = ps.instr_hi > ps.instr_lo && mx >= 0;
            lo_min[gp] = nonempty[gp] ? mn : P.G;
            hi_max[gp] = nonempty[gp] ? mx + 1 : 0;
        }
    }
    std::vector<int> suf(np + 1, P.G), pre(np + 1, 0);         // min instruction of passes >= p, 1 + max instruction of passes < p
    for (int p = np - 1; p >= 0; --p) suf[p] = std::min(suf[p + 1], lo_min[p]);
    for (int p = 0; p < np; ++p) pre[p + 1] = std::max(pre[p], hi_max[p]);
    std::vector<int32_t> out(2 * (size_t)(P.G + 1));
    int best_pass = -1, best_L = 0, p = 0;
    for (int e = 0; e <= P.G; ++e) {
        for (; p < np && suf[p] <= e; ++p)
            if (nonempty[p] && pre[p] <= suf[p]) { best_pass = p; best_L = suf[p]; }
        if (e == P.G) { best_pass = np; best_L = P.G; }
        out[2 * e] = best_pass; out[2 * e + 1] = best_L;
    }
    return out;
}

std::string spec_resident_source(const Program& P, const SpecOptions& opt, int which) {
    std::ostringstream o, k;
    o << "// generated by ncb_spec.cpp -- do not edit\n#include \"ncb_spec.cuh\"\nusing namespace ncb;\n";
    if (!spec_resident_supported(P) || !emit_resident_module(k, P, opt, which)) return "";
    o << k.str();
    return o.str();
}

uint64_t spec_structure_hash(const Program& P, const SpecOptions& opt) {
    // everything the generators read, without generating: geometry, the segments' blobs up to their pools (headers, passes,
    // operator records with their table offsets), the pools' sizes, the constant-bank programs (with the coefficients only
    // when they are written out as literals)
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* ptr, size_t n) {
        const unsigned char* b = (const unsigned char*)ptr;
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    };
    const int geo[8] = {P.nbits, P.K, P.R, P.G, P.mode, (int)P.segs.size(), (int)P.passes.size(), opt.literal};
    mix(geo, sizeof geo);
    const int o2[7] = {opt.nbuf, opt.load, opt.events, opt.factor, opt.ctas, opt.resident_fast, opt.ev_fast};
    mix(o2, sizeof o2);
    for (size_t sgi = 0; sgi < P.blob_off.size(); ++sgi) {
        const unsigned char* braw = P.blobs.data() + P.blob_off[sgi];
        const SegBlob& blob = *reinterpret_cast<const SegBlob*>(braw);
        mix(braw, (size_t)blob.pool_off);
        mix(&blob.bytes, sizeof blob.bytes);
    }
    const bool tiled = lean_tiled(P);
    for (size_t sgi = 0; tiled && sgi < P.segconst.size(); ++sgi)
        mix(&P.segconst[sgi], opt.literal ? sizeof(SegConst) : offsetof(SegConst, coef));
    return h;
}

std::string spec_key(const Program& P, const SpecOptions& opt) {
    // FNV-1a over the generated source of every segment: the source is all the cubins depend on (beside the headers,
    // which the engine covers with a build stamp)
    uint64_t h = 1469598103934665603ull;
    const std::string src = spec_source(P, 0, (int)P.segs.size(), opt) + spec_resident_source(P, opt);
    for (unsigned char ch : src) { h ^= ch; h *= 1099511628211ull; }
    const int geo[4] = {P.nbits, P.K, P.R, 17 /* generator version */};
    for (size_t i = 0; i < sizeof geo; ++i) { h ^= reinterpret_cast<const unsigned char*>(geo)[i]; h *= 1099511628211ull; }
    char b[32];
    std::snprintf(b, sizeof b, "%016llx", (unsigned long long)h);
    return b;
}

}  // namespace ncb
