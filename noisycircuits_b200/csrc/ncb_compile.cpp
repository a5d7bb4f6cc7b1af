// Host-side compiler of the B200 noisy-circuit engine (see ncb_compile.h).
#include <memory>
#include <mutex>
#include "ncb_compile.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <sstream>
#include <stdexcept>

namespace ncb {

// Multiply-gather of <= 3 tile-index bits (positions `pos`, each < 16): finds `magic` such that
// h(x) = ((x & mask) * magic) >> 29 (32-bit arithmetic) maps the 2^n settings of those bits one-to-one onto [0, 2^n).
// The search space is small (every position set below 16 bits resolves within a few hundred candidates).
struct Gather { uint32_t mask, magic; };
static Gather find_gather(const std::vector<int>& pos) {
    const int n = (int)pos.size();
    Gather g{0u, 0u};
    if (n == 0) return g;
    if (n > 3) throw std::runtime_error("find_gather: more than 3 bits");
    uint32_t keys[8];
    for (int s = 0; s < (1 << n); ++s) {
        keys[s] = 0;
        for (int i = 0; i < n; ++i) {
            if (pos[i] < 0 || pos[i] > 15) throw std::runtime_error("find_gather: bit position out of range");
            keys[s] |= (uint32_t)((s >> i) & 1) << pos[i];
        }
    }
    g.mask = keys[(1 << n) - 1];
    uint64_t st = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 16); };
    for (long long t = 0; t < 50000000ll; ++t) {
        uint32_t m = 0;
        if (t == 0) { for (int i = 0; i < n; ++i) m += 1u << (29 + i - pos[i]); }
        else if (t == 1) { for (int i = 0; i < n; ++i) m += 1u << (29 + (n - 1 - i) - pos[i]); }
        else { m = rnd(); if (t % 3) m &= rnd(); if (t % 3 == 2) m &= rnd(); }
        uint32_t seen = 0;
        bool ok = true;
        for (int s = 0; s < (1 << n) && ok; ++s) {
            const uint32_t h = (keys[s] * m) >> 29;
            if (h >= (1u << n) || ((seen >> h) & 1u)) ok = false;
            seen |= 1u << h;
        }
        if (ok) { g.magic = m; return g; }
    }
    throw std::runtime_error("find_gather: no multiplier found");
}
namespace {

using cd = std::complex<double>;
using cmat = std::vector<cd>;   // row-major square matrix

cmat matmul(const cmat& a, const cmat& b, int d) {
    cmat r(d * d, cd(0, 0));
    for (int i = 0; i < d; ++i)
        for (int k = 0; k < d; ++k) {
            const cd x = a[i * d + k];
            if (x == cd(0, 0)) continue;
            for (int j = 0; j < d; ++j) r[i * d + j] += x * b[k * d + j];
        }
    return r;
}
cmat adjoint_times(const cmat& a, int d) {   // a^+ a
    cmat r(d * d, cd(0, 0));
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            cd s(0, 0);
            for (int k = 0; k < d; ++k) s += std::conj(a[k * d + i]) * a[k * d + j];
            r[i * d + j] = s;
        }
    return r;
}
double max_abs(const cmat& a) {
    double m = 0;
    for (const cd& x : a) m = std::max(m, std::abs(x));
    return m;
}
bool is_diagonal(const cmat& a, int d) {
    const double tol = 1e-17 * std::max(1.0, max_abs(a));
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j)
            if (i != j && std::abs(a[i * d + j]) > tol) return false;
    return true;
}

// Extreme eigenvalues of a Hermitian d x d matrix (d <= 4): cyclic Jacobi on the real symmetric embedding
// [[A, -B], [B, A]] of H = A + iB (its spectrum is that of H, doubled).
void herm_eig_minmax(const cmat& h, int d, double* emin, double* emax) {
    const int n = 2 * d;
    double s[8][8];
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            const cd x = 0.5 * (h[i * d + j] + std::conj(h[j * d + i]));
            s[i][j] = x.real(); s[i + d][j + d] = x.real();
            s[i][j + d] = -x.imag(); s[i + d][j] = x.imag();
        }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) off += s[p][q] * s[p][q];
        if (off < 1e-40) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                if (std::fabs(s[p][q]) < 1e-300) continue;
                const double theta = (s[q][q] - s[p][p]) / (2.0 * s[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < n; ++k) {
                    const double a = s[k][p], b = s[k][q];
                    s[k][p] = c * a - sn * b; s[k][q] = sn * a + c * b;
                }
                for (int k = 0; k < n; ++k) {
                    const double a = s[p][k], b = s[q][k];
                    s[p][k] = c * a - sn * b; s[q][k] = sn * a + c * b;
                }
            }
    }
    double lo = s[0][0], hi = s[0][0];
    for (int i = 1; i < n; ++i) { lo = std::min(lo, s[i][i]); hi = std::max(hi, s[i][i]); }
    *emin = lo; *emax = hi;
}

struct PoolBuilder {
    std::vector<double>& pool;
    int add_complex(const cmat& m) {
        if (pool.size() & 1) pool.push_back(0.0);
        const int off = (int)pool.size();
        for (const cd& x : m) { pool.push_back(x.real()); pool.push_back(x.imag()); }
        return off;
    }
    int add_doubles(const std::vector<double>& v) {
        if (pool.size() & 1) pool.push_back(0.0);
        const int off = (int)pool.size();
        pool.insert(pool.end(), v.begin(), v.end());
        if (pool.size() & 1) pool.push_back(0.0);
        return off;
    }
};

cmat read_complex(const double* p, int count) {
    cmat m(count);
    for (int i = 0; i < count; ++i) m[i] = cd(p[2 * i], p[2 * i + 1]);
    return m;
}
std::vector<double> flatten(const cmat& m) {
    std::vector<double> v;
    v.reserve(2 * m.size());
    for (const cd& x : m) { v.push_back(x.real()); v.push_back(x.imag()); }
    return v;
}
cmat diag_of(const cmat& m, int d) {
    cmat v(d);
    for (int i = 0; i < d; ++i) v[i] = m[i * d + i];
    return v;
}

struct ChannelInfo {
    std::vector<cmat> raw;      // K_k
    std::vector<cmat> scaled;   // s_k K_k
    std::vector<double> scale;  // s_k
    std::vector<bool> unitary_like;
    bool certain_nonunitary = false;
};

}  // namespace

void gate_matrix(int opcode, double theta, double* out, int* dim) {
    const double c = std::cos(0.5 * theta), s = std::sin(0.5 * theta);
    const double r2 = 0.70710678118654752440;
    const cd I(0, 1);
    cmat m;
    switch (opcode) {
    case G_X: m = {0, 1, 1, 0}; break;
    case G_SX: m = {cd(0.5, 0.5), cd(0.5, -0.5), cd(0.5, -0.5), cd(0.5, 0.5)}; break;
    case G_RZ: m = {cd(c, -s), 0, 0, cd(c, s)}; break;
    case G_RX: m = {c, -I * s, -I * s, c}; break;
    case G_RY: m = {c, -s, s, c}; break;
    case G_H: m = {r2, r2, r2, -r2}; break;
    case G_P: m = {1, 0, 0, cd(std::cos(theta), std::sin(theta))}; break;   // QuantumGates.hpp:689-698
    case G_CZ: m = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1}; break;
    case G_RZZ: m = {cd(c, -s), 0, 0, 0, 0, cd(c, s), 0, 0, 0, 0, cd(c, s), 0, 0, 0, 0, cd(c, -s)}; break;
    case G_ECR: m = {0, r2, 0, I * r2, r2, 0, -I * r2, 0, 0, I * r2, 0, r2, -I * r2, 0, r2, 0}; break;
    case G_CX: m = {1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0}; break;      // control = first listed qubit
    case G_SWAP: m = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1}; break;
    default: throw std::runtime_error("gate_matrix: unknown opcode " + std::to_string(opcode));
    }
    *dim = m.size() == 4 ? 2 : 4;
    for (size_t i = 0; i < m.size(); ++i) { out[2 * i] = m[i].real(); out[2 * i + 1] = m[i].imag(); }
}

namespace {

// Kraus list -> Channel record (+ pool data) and ChannelInfo for merging.
Channel analyse_channel(const CircuitView& c, int ch, PoolBuilder& pb, ChannelInfo& info) {
    const int d = c.channel_dim[ch], count = c.channel_count[ch], dd = d * d;
    if (d != 2 && d != 4) throw std::runtime_error("channel " + std::to_string(ch) + ": operator dimension must be 2 or 4");
    if (count < 1) throw std::runtime_error("channel " + std::to_string(ch) + ": empty Kraus list");
    const double* src = c.kraus_pool + 2 * c.channel_offset[ch];
    std::vector<cmat> mk(count);
    std::vector<double> wmin(count), wmax(count);
    info.raw.resize(count); info.scaled.resize(count); info.scale.resize(count); info.unitary_like.resize(count);
    for (int k = 0; k < count; ++k) {
        info.raw[k] = read_complex(src + 2 * (size_t)k * dd, dd);
        mk[k] = adjoint_times(info.raw[k], d);
        herm_eig_minmax(mk[k], d, &wmin[k], &wmax[k]);
        const double w = 0.5 * (wmin[k] + wmax[k]);
        info.unitary_like[k] = (wmax[k] - wmin[k]) <= 1e-13 * std::max(wmax[k], 1e-300) && wmax[k] > 0;
        const double ref = info.unitary_like[k] ? w : wmax[k];
        info.scale[k] = ref > 0 ? 1.0 / std::sqrt(ref) : 1.0;
        info.scaled[k] = info.raw[k];
        for (cd& x : info.scaled[k]) x *= info.scale[k];
    }
    // cumulative sums and state-independent bounds of cp[k] = <C_k>/<C_total>
    std::vector<double> lo(std::max(count - 1, 1), 1.0), hi(std::max(count - 1, 1), 1.0);
    cmat total(dd, cd(0, 0));
    for (int k = 0; k < count; ++k)
        for (int i = 0; i < dd; ++i) total[i] += mk[k][i];
    double tmin, tmax;
    herm_eig_minmax(total, d, &tmin, &tmax);
    if (!(tmax > 0)) throw std::runtime_error("channel " + std::to_string(ch) + ": all Kraus operators are zero");
    cmat cum(dd, cd(0, 0));
    for (int k = 0; k + 1 < count; ++k) {
        for (int i = 0; i < dd; ++i) cum[i] += mk[k][i];
        double cmin, cmax;
        herm_eig_minmax(cum, d, &cmin, &cmax);
        cmin = std::max(cmin, 0.0); cmax = std::max(cmax, 0.0);
        lo[k] = std::min(1.0, cmin / tmax);
        hi[k] = tmin > 0 ? std::min(1.0, cmax / tmin) : 1.0;
        if (k > 0) { lo[k] = std::max(lo[k], lo[k - 1]); hi[k] = std::max(hi[k], hi[k - 1]); }
    }
    // base branch = widest state-independent interval
    int base = 0;
    double best = -1;
    for (int k = 0; k < count; ++k) {
        const double left = k == 0 ? 0.0 : hi[k - 1], right = k == count - 1 ? 1.0 : lo[k];
        const double w = right - left;
        if (w > 2 * EVENT_MARGIN && !info.unitary_like[k]) info.certain_nonunitary = true;
        if (w > best) { best = w; base = k; }
    }
    if (count == 1 && !info.unitary_like[0]) info.certain_nonunitary = true;
    Channel out;
    std::memset(&out, 0, sizeof(out));
    out.dim = d; out.count = count; out.base = base;
    cmat all_scaled, all_mk;
    for (int k = 0; k < count; ++k) {
        all_scaled.insert(all_scaled.end(), info.scaled[k].begin(), info.scaled[k].end());
        all_mk.insert(all_mk.end(), mk[k].begin(), mk[k].end());
    }
    out.kraus = pb.add_complex(all_scaled);
    out.mk = pb.add_complex(all_mk);
    std::vector<double> bounds(lo.begin(), lo.begin() + std::max(count - 1, 0));
    bounds.insert(bounds.end(), hi.begin(), hi.begin() + std::max(count - 1, 0));
    if (bounds.empty()) bounds.push_back(1.0);
    out.bounds = pb.add_doubles(bounds);
    out.scale = pb.add_doubles(info.scale);
    return out;
}

// superoperator sum_k A_k (x) conj(A_k) with vectorised index = row + d*col
cmat superop(const std::vector<cmat>& A, int d) {
    const int D = d * d;
    cmat s(D * D, cd(0, 0));
    for (const cmat& a : A) {
        if (max_abs(a) == 0) continue;
        for (int r2 = 0; r2 < d; ++r2)
            for (int c2 = 0; c2 < d; ++c2)
                for (int r = 0; r < d; ++r)
                    for (int c = 0; c < d; ++c) s[(r2 + d * c2) * D + (r + d * c)] += a[r2 * d + r] * std::conj(a[c2 * d + c]);
    }
    return s;
}

LogicalOp make_lop(int instr, const std::vector<int>& bits, const cmat& merged, const cmat& bare) {
    LogicalOp l;
    l.instr = instr;
    l.nb = (int)bits.size();
    for (int i = 0; i < 8; ++i) l.bits[i] = i < l.nb ? bits[i] : -1;
    const int d = 1 << l.nb;
    l.generic = l.nb > 4;
    l.diag = !l.generic && is_diagonal(merged, d) && is_diagonal(bare, d);
    if (l.diag) { l.mat = flatten(diag_of(merged, d)); l.bare = flatten(diag_of(bare, d)); }
    else { l.mat = flatten(merged); l.bare = flatten(bare); }
    return l;
}

// One-qubit operator patterns the executor has cheap kernels for (ncb_common.h: OPK_*).
int pattern_1q(const cmat& m, double tol) {
    const double s = std::max(1.0, max_abs(m));
    auto z = [&](double x) { return std::fabs(x) <= tol * s; };
    if (z(m[0].imag()) && z(m[3].imag()) && z(m[1].real()) && z(m[2].real())) return OPK_RX1;
    return OPK_DENSE1;
}
void snap_pattern(cmat& m, int kind) {      // remove the rounding dust the pattern ignores
    if (kind == OPK_RX1) { m[0] = m[0].real(); m[3] = m[3].real(); m[1] = cd(0, m[1].imag()); m[2] = cd(0, m[2].imag()); }
}
// Factors a global phase e^{i phi} out of a one-qubit gate (bare and merged with its base Kraus operator alike) when
// that turns it into the cheap pattern (sx = e^{i pi/4} rx(pi/2), x = i rx(pi)).  Returns phi (0: nothing factored).
double dephase_1q(cmat& merged, cmat& bare) {
    const double tol = 4e-16;
    if (pattern_1q(bare, tol) == OPK_RX1 && pattern_1q(merged, tol) == OPK_RX1) return 0.0;
    std::vector<double> cands;
    for (int i = 0; i < 4; ++i)
        if (std::abs(bare[i]) > 1e-3) cands.push_back(std::arg(bare[i]) - ((i == 1 || i == 2) ? M_PI / 2 : 0.0));
    for (double phi : cands) {
        const cd f = std::polar(1.0, -phi);
        cmat b2(bare), m2(merged);
        for (cd& x : b2) x *= f;
        for (cd& x : m2) x *= f;
        if (pattern_1q(b2, tol) == OPK_RX1 && pattern_1q(m2, tol) == OPK_RX1) { bare = b2; merged = m2; return phi; }
    }
    return 0.0;
}

// One-qubit operator M as D1 . X . D2 with D1, D2 diagonal phases and X of the cheap pattern [[a, i b], [i c, d]] (a, b, c, d
// real): possible when M00 M11 / (M01 M10) is real -- every unitary, and a unitary times a diagonal Kraus operator.  The
// phases then move into neighbouring phase tables (diagonal operators commute) and what stays costs half (a quarter with one
// coefficient factored out) of the generic 2x2.  Gauge: D2 = diag(1, e^{i psi}).  Returns false when M has no such form.
bool split_dxd(const cmat& M, cd d1[2], cmat& X, cd d2[2]) {
    const double s = max_abs(M);
    if (!(s > 0.0)) return false;
    const double tiny = 1e-14 * s;
    auto ph = [&](const cd& z) { return std::abs(z) > tiny ? z / std::abs(z) : cd(1, 0); };
    const cd I(0, 1);
    cd f0, f1, g1;                    // e^{i phi0}, e^{i phi1}, e^{i psi1}; e^{i psi0} = 1
    const bool z00 = std::abs(M[0]) <= tiny, z01 = std::abs(M[1]) <= tiny, z10 = std::abs(M[2]) <= tiny, z11 = std::abs(M[3]) <= tiny;
    if (!z00) {
        f0 = ph(M[0]);
        g1 = !z01 ? ph(M[1]) / (I * f0) : cd(1, 0);
        if (!z10) f1 = ph(M[2]) / I;
        else f1 = !z11 ? ph(M[3]) / g1 : cd(1, 0);
        if (z01 && !z10 && !z11) g1 = ph(M[3]) / f1;
    } else {
        // M00 = 0: phi0 is free unless M01 pins phi0 + psi1
        f1 = !z10 ? ph(M[2]) / I : cd(1, 0);
        g1 = !z11 ? ph(M[3]) / f1 : cd(1, 0);
        f0 = !z01 ? ph(M[1]) / (I * g1) : cd(1, 0);
    }
    d1[0] = f0; d1[1] = f1; d2[0] = cd(1, 0); d2[1] = g1;
    X.assign(4, cd(0, 0));
    X[0] = M[0] / (d1[0] * d2[0]); X[1] = M[1] / (d1[0] * d2[1]); X[2] = M[2] / (d1[1] * d2[0]); X[3] = M[3] / (d1[1] * d2[1]);
    if (pattern_1q(X, 4e-16) != OPK_RX1) return false;
    snap_pattern(X, OPK_RX1);
    // verify: D1 X D2 == M to rounding
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c)
            if (std::abs(d1[r] * X[r * 2 + c] * d2[c] - M[r * 2 + c]) > 4e-16 * std::max(1.0, s)) return false;
    return true;
}

// Two-qubit gate U (index = b0 + 2 b1) as (A1 x A0) . D . (B1 x B0) with D diagonal, for gates that map each basis
// state of one "control" bit to a basis state of it (controlled-V, ecr, cx, cy, ...):
//   U = P_c . sum_b |b><b|_c (x) V_b,   V_0^+ V_1 = Q L Q^+   =>   A_c = P_c (1 or X), A_t = V_0 Q, B_t = Q^+, B_c = 1,
//   D = diag over (c, t): 1 for c = 0, L_t for c = 1.
// The dense 4x4 then runs as one-qubit operators (which merge with their neighbours) and a diagonal table in the lean
// kernels.  Returns false when U has no such form (swap, generic unitaries): those stay dense.
struct ControlledForm { cmat A[2], B[2], D; };
bool controlled_form(const cmat& U, ControlledForm& f) {
    const double tol = 1e-13;
    for (int c = 0; c < 2; ++c) {
        const int t = 1 - c;
        auto at = [&](int cr, int tr, int cc, int tc) { return U[(size_t)((cr << c) | (tr << t)) * 4 + ((cc << c) | (tc << t))]; };
        int pi[2];
        bool ok = true;
        for (int cc = 0; cc < 2 && ok; ++cc) {
            double w[2] = {0, 0};
            for (int cr = 0; cr < 2; ++cr)
                for (int tr = 0; tr < 2; ++tr)
                    for (int tc = 0; tc < 2; ++tc) w[cr] = std::max(w[cr], std::abs(at(cr, tr, cc, tc)));
            if (w[0] > tol && w[1] > tol) ok = false;
            pi[cc] = w[1] > w[0] ? 1 : 0;
        }
        if (!ok || pi[0] == pi[1]) continue;
        cmat V[2];
        for (int cc = 0; cc < 2; ++cc) V[cc] = {at(pi[cc], 0, cc, 0), at(pi[cc], 0, cc, 1), at(pi[cc], 1, cc, 0), at(pi[cc], 1, cc, 1)};
        auto dagger = [](const cmat& m) { return cmat{std::conj(m[0]), std::conj(m[2]), std::conj(m[1]), std::conj(m[3])}; };
        const cmat W = matmul(dagger(V[0]), V[1], 2);
        // eigen-decomposition of the 2x2 unitary W
        cmat Q{1, 0, 0, 1};
        cd l0 = W[0], l1 = W[3];
        if (std::abs(W[1]) > 1e-15 || std::abs(W[2]) > 1e-15) {
            const cd tr = W[0] + W[3], det = W[0] * W[3] - W[1] * W[2];
            const cd disc = std::sqrt(tr * tr - 4.0 * det);
            l0 = 0.5 * (tr + disc);
            cd v0, v1;
            if (std::abs(W[1]) >= std::abs(W[2])) { v0 = W[1]; v1 = l0 - W[0]; } else { v0 = l0 - W[3]; v1 = W[2]; }
            const double nv = std::sqrt(std::norm(v0) + std::norm(v1));
            if (nv < 1e-14) continue;
            v0 /= nv; v1 /= nv;
            Q = {v0, -std::conj(v1), v1, std::conj(v0)};       // columns: eigenvector, its orthogonal complement
            const cmat L = matmul(dagger(Q), matmul(W, Q, 2), 2);
            if (std::abs(L[1]) > tol || std::abs(L[2]) > tol) continue;
            l0 = L[0]; l1 = L[3];
        }
        const cmat one{1, 0, 0, 1}, X{0, 1, 1, 0};
        f.A[c] = pi[0] == 0 ? one : X;
        f.A[t] = matmul(V[0], Q, 2);
        f.B[c] = one;
        f.B[t] = dagger(Q);
        f.D.assign(4, cd(1, 0));
        for (int tb = 0; tb < 2; ++tb) f.D[(size_t)((1 << c) | (tb << t))] = tb ? l1 : l0;
        // verify
        auto kron = [](const cmat& m1, const cmat& m0) {     // m1 on bit 1, m0 on bit 0
            cmat k(16);
            for (int r = 0; r < 4; ++r)
                for (int cc = 0; cc < 4; ++cc) k[(size_t)r * 4 + cc] = m1[(size_t)(r >> 1) * 2 + (cc >> 1)] * m0[(size_t)(r & 1) * 2 + (cc & 1)];
            return k;
        };
        cmat Dm(16, cd(0, 0));
        for (int i = 0; i < 4; ++i) Dm[(size_t)i * 5] = f.D[i];
        const cmat R = matmul(kron(f.A[1], f.A[0]), matmul(Dm, kron(f.B[1], f.B[0]), 4), 4);
        double err = 0;
        for (int i = 0; i < 16; ++i) err = std::max(err, std::abs(R[i] - U[i]));
        if (err < 1e-13) return true;
    }
    return false;
}

cmat conj_of(const cmat& a) {
    cmat r(a);
    for (cd& x : r) x = std::conj(x);
    return r;
}

// deposit "source bit i -> output bit pos[i]" (i < nbits) as runs of consecutive source bits with consecutive targets
BitRuns make_runs(const std::vector<int>& pos) {
    BitRuns r;
    std::memset(&r, 0, sizeof(r));
    const int nbits = (int)pos.size();
    int i = 0;
    while (i < nbits) {
        int j = i + 1;
        while (j < nbits && pos[j] == pos[j - 1] + 1) ++j;
        if (r.n >= MAX_RUNS) throw std::runtime_error("internal: too many bit runs");
        r.mask[r.n] = (uint32_t)(((1ull << j) - 1) ^ ((1ull << i) - 1));
        r.shift[r.n] = (int8_t)(pos[i] - i);
        ++r.n;
        i = j;
    }
    return r;
}

struct PassBuild {
    std::vector<int> regset;            // tile-local positions
    std::vector<const LogicalOp*> lops;
};

}  // namespace

static void build_blobs(Program& P);
static void compile_program_once(const CircuitView& c, int mode, const CompileOptions& opt, Program& P);

namespace { struct BlobOverflow : std::runtime_error { using std::runtime_error::runtime_error; }; }

// The segmenter cuts segments with an ESTIMATE of their staged program size; should a segment still come out larger
// than the staging area, the program is compiled again with a more cautious estimate.
void compile_program(const CircuitView& c, int mode, const CompileOptions& opt, Program& P) {
    CompileOptions o = opt;
    if (const char* e = getenv("NCB_BLOB_COST_SCALE")) o.blob_cost_scale = std::max(1e-3, atof(e));      // test hook
    for (int attempt = 0;; ++attempt) {
        try {
            compile_program_once(c, mode, o, P);
            P.opt.blob_cost_scale = opt.blob_cost_scale;
            P.compile_attempts = attempt + 1;
            return;
        } catch (const BlobOverflow&) {
            if (attempt >= 12) throw;
            o.blob_cost_scale *= 1.6;
        }
    }
}

static void compile_program_once(const CircuitView& c, int mode, const CompileOptions& opt, Program& P) {
    P = Program();
    P.mode = mode; P.opt = opt;
    P.nq = c.num_qubits; P.G = c.num_instructions;
    if (P.nq < 1 || P.nq > MAX_QUBITS) throw std::runtime_error("num_qubits out of range");
    P.nbits = mode == MODE_DENSITY ? 2 * P.nq : P.nq;
    if (P.nbits > MAX_QUBITS) throw std::runtime_error("state too large: " + std::to_string(P.nbits) + " state bits");
    P.R = std::min(std::max(opt.reg_bits, 1), MAX_REG_BITS);
    const int kmin = P.R + 5;
    P.K = std::min(std::max({opt.tile_bits, kmin}), MAX_TILE_BITS);
    if (P.nbits <= std::max(P.K, std::min(opt.resident_bits, MAX_TILE_BITS - 1)))
        P.K = std::max(P.nbits, kmin);                           // whole state in one tile (extra bits are virtual)
    const int K = P.K, R = P.R;
    const int L = std::min(std::max(opt.low_bits, 0), std::min(K - 4, P.nbits));
    PoolBuilder pb{P.pool};

    // ---- channels -------------------------------------------------------------------------------------------
    std::vector<ChannelInfo> cinfo_own(std::max(c.num_channels, 0));
    std::shared_ptr<const void> cinfo_keep;                 // (a cached analysis stays alive while it is read)
    const std::vector<ChannelInfo>* cinfo_p = &cinfo_own;
    const bool use_noise = mode != MODE_PURE;
    if (use_noise && c.num_channels > 0) {
        // The analysis of the Kraus lists (operators, M_k, eigenvalue bounds of the cumulative sums) depends on the noise model
        // only: a parameter sweep re-compiles one gate skeleton with the same channels thousands of times (ncb_program_update),
        // so the last few analysed channel sets are kept, keyed by the channel data itself.
        struct Entry { std::vector<unsigned char> key; std::vector<Channel> chans; std::vector<ChannelInfo> info; std::vector<double> pool; };
        static std::mutex mu;
        static std::vector<std::shared_ptr<const Entry>> cache;
        std::vector<unsigned char> key;
        {
            const int nch = c.num_channels;
            const long long elems = c.channel_offset[nch - 1] + (long long)c.channel_count[nch - 1] * c.channel_dim[nch - 1] * c.channel_dim[nch - 1];
            auto put = [&](const void* ptr, size_t n) { const unsigned char* b = (const unsigned char*)ptr; key.insert(key.end(), b, b + n); };
            put(&nch, sizeof nch);
            put(c.channel_offset, sizeof(int64_t) * nch); put(c.channel_count, sizeof(int32_t) * nch); put(c.channel_dim, sizeof(int32_t) * nch);
            put(c.kraus_pool, sizeof(double) * 2 * (size_t)elems);
        }
        std::shared_ptr<const Entry> hit;
        {
            std::lock_guard<std::mutex> lk(mu);
            for (const auto& e : cache) if (e->key == key) { hit = e; break; }
        }
        if (!hit) {
            auto e = std::make_shared<Entry>();
            for (int ch = 0; ch < c.num_channels; ++ch) P.chans.push_back(analyse_channel(c, ch, pb, cinfo_own[ch]));
            e->key = std::move(key); e->chans = P.chans; e->info = cinfo_own; e->pool = P.pool;
            std::lock_guard<std::mutex> lk(mu);
            cache.insert(cache.begin(), e);
            if (cache.size() > 4) cache.pop_back();
        } else {
            P.chans = hit->chans; cinfo_p = &hit->info; cinfo_keep = hit; P.pool = hit->pool;
        }
    }
    const std::vector<ChannelInfo>& cinfo = *cinfo_p;

    // ---- logical ops ----------------------------------------------------------------------------------------
    std::vector<LogicalOp> lops;
    P.instrs.resize(P.G);
    P.draw_pos.assign(P.G, -1);
    for (int g = 0; g < P.G; ++g) {
        const int opc = c.opcode[g];
        Instr& in = P.instrs[g];
        in.channel = -1; in.q0 = c.q0[g]; in.q1 = -1; in.seg = -1;
        std::vector<int> qb;
        cmat U;
        if (opc == G_UNITARY) {
            const int k = c.unitary_nq[g];
            if (k < 1 || k > 8) throw std::runtime_error("instruction " + std::to_string(g) + ": unitary on " + std::to_string(k) + " qubits is not supported (1..8)");
            for (int b = 0; b < k; ++b) qb.push_back(c.unitary_qubits[8 * g + b]);
            U = read_complex(c.unitary_pool + 2 * c.unitary_offset[g], (1 << k) * (1 << k));
        } else {
            if (opc < 0 || opc >= G_UNITARY) throw std::runtime_error("instruction " + std::to_string(g) + ": unknown opcode " + std::to_string(opc));
            double buf[32]; int d;
            gate_matrix(opc, c.theta[g], buf, &d);
            U = read_complex(buf, d * d);
            qb.push_back(c.q0[g]);
            if (d == 4) { qb.push_back(c.q1[g]); in.q1 = c.q1[g]; }
        }
        for (size_t a = 0; a < qb.size(); ++a) {
            if (qb[a] < 0 || qb[a] >= P.nq) throw std::runtime_error("instruction " + std::to_string(g) + ": qubit out of range");
            for (size_t b = a + 1; b < qb.size(); ++b)
                if (qb[a] == qb[b]) throw std::runtime_error("instruction " + std::to_string(g) + ": repeated qubit");
        }
        const int d = 1 << qb.size();
        int ch = (use_noise && opc != G_UNITARY && c.channel) ? c.channel[g] : -1;
        if (ch >= c.num_channels) throw std::runtime_error("instruction " + std::to_string(g) + ": channel id out of range");
        if (ch >= 0 && P.chans[ch].dim != d) throw std::runtime_error("instruction " + std::to_string(g) + ": channel dimension does not match the gate");
        if (mode == MODE_DENSITY) {
            std::vector<int> cb;
            for (int q : qb) cb.push_back(q + P.nq);
            if (ch >= 0) {
                std::vector<cmat> A;
                for (const cmat& k : cinfo[ch].raw) A.push_back(matmul(k, U, d));
                std::vector<int> bits(qb);
                bits.insert(bits.end(), cb.begin(), cb.end());
                const cmat S = superop(A, d);
                lops.push_back(make_lop(g, bits, S, S));
            } else {
                lops.push_back(make_lop(g, qb, U, U));
                const cmat Uc = conj_of(U);
                lops.push_back(make_lop(g, cb, Uc, Uc));
            }
            continue;
        }
        cmat merged = U;
        if (ch >= 0) {
            in.channel = ch;
            merged = matmul(cinfo[ch].scaled[P.chans[ch].base], U, d);
            if (cinfo[ch].certain_nonunitary) P.norm_preserving = false;
            if (P.chans[ch].count >= 2) { P.draw_pos[g] = P.num_draws++; P.noisy_instrs.push_back(g); }
        }
        auto push_1q = [&](int q, cmat m, cmat b) {     // one-qubit lop with its global phase factored out
            if (!(is_diagonal(m, 2) && is_diagonal(b, 2))) {
                const double phi = dephase_1q(m, b);
                if (phi != 0.0) {
                    const cd gp = cd(P.global_phase[0], P.global_phase[1]) * std::polar(1.0, phi);
                    P.global_phase[0] = gp.real(); P.global_phase[1] = gp.imag();
                }
            }
            lops.push_back(make_lop(g, std::vector<int>{q}, m, b));
        };
        if (qb.size() == 1) { push_1q(qb[0], merged, U); continue; }
        // controlled-type two-qubit gates (ecr, cx): one-qubit factors + a diagonal table instead of a dense 4x4, when
        // the channel's base operator is diagonal too (it then is one more diagonal lop whose bare form is the identity)
        ControlledForm cf;
        const bool base_diag = ch < 0 || is_diagonal(cinfo[ch].scaled[P.chans[ch].base], 4);
        if (qb.size() == 2 && opt.fuse && opt.decompose && !is_diagonal(U, 4) && base_diag && controlled_form(U, cf)) {
            auto is_one = [](const cmat& m) { return m[0] == cd(1, 0) && m[3] == cd(1, 0) && m[1] == cd(0, 0) && m[2] == cd(0, 0); };
            for (int b = 0; b < 2; ++b)
                if (!is_one(cf.B[b])) push_1q(qb[b], cf.B[b], cf.B[b]);
            cmat Dm(16, cd(0, 0)), one4(16, cd(0, 0));
            for (int i = 0; i < 4; ++i) { Dm[(size_t)i * 5] = cf.D[i]; one4[(size_t)i * 5] = 1; }
            lops.push_back(make_lop(g, qb, Dm, Dm));
            for (int b = 0; b < 2; ++b)
                if (!is_one(cf.A[b])) push_1q(qb[b], cf.A[b], cf.A[b]);
            if (ch >= 0) lops.push_back(make_lop(g, qb, cinfo[ch].scaled[P.chans[ch].base], one4));
            continue;
        }
        lops.push_back(make_lop(g, qb, merged, U));
    }

    const int kfit = std::min(K, P.nbits);
    // ---- execution order ------------------------------------------------------------------------------------------
    // Instructions on disjoint bits commute, so any order that keeps the per-bit order is the same circuit.  The
    // scheduler below picks, segment by segment, a tile (the L low bits + up to kfit - L others) and runs every
    // instruction that is ready inside it -- several layers deep where the coupling allows (space-time trapezoids) --
    // which is what decides how many times the state crosses HBM.  Jump events keep the reference's meaning on the
    // order chosen here: the engine is the reference algorithm applied to this (equivalent) instruction order, and
    // `orig` lets a caller replay exactly that.
    {
        std::vector<std::vector<int>> ilops(P.G);          // lops of each instruction
        std::vector<uint64_t> ubits(P.G, 0);
        for (size_t i = 0; i < lops.size(); ++i) {
            ilops[lops[i].instr].push_back((int)i);
            for (int b = 0; b < lops[i].nb; ++b) ubits[lops[i].instr] |= 1ull << lops[i].bits[b];
        }
        std::vector<int> order;
        order.reserve(P.G);
        if (!opt.reorder || !opt.fuse || P.nbits <= kfit) {
            for (int g = 0; g < P.G; ++g) order.push_back(g);
        } else {
            const int nb = P.nbits;
            std::vector<std::vector<int>> queue(nb);
            for (int g = 0; g < P.G; ++g)
                for (int b = 0; b < nb; ++b)
                    if ((ubits[g] >> b) & 1) queue[b].push_back(g);
            std::vector<size_t> head(nb, 0);
            uint64_t low = 0;
            for (int b = 0; b < L; ++b) low |= 1ull << b;
            int max_seg_ops = 50;        // keeps a segment's program inside the shared-memory staging area
            if (const char* e = getenv("NCB_MAX_SEG_OPS")) max_seg_ops = std::max(1, atoi(e));
            std::vector<char> generic_ptr(P.G, 0);
            for (const LogicalOp& l : lops) if (l.generic) generic_ptr[l.instr] = 1;
            // runs every ready instruction inside tile S from heads h (modified); returns the executed list
            auto run = [&](uint64_t S, std::vector<size_t>& h, std::vector<int>* out) {
                int count = 0;
                bool progress = true;
                std::vector<int> ready;
                while (progress && count < max_seg_ops) {
                    progress = false;
                    ready.clear();
                    for (int b = 0; b < nb; ++b) {
                        if (!((S >> b) & 1) || h[b] >= queue[b].size()) continue;
                        const int g = queue[b][h[b]];
                        if (ubits[g] & ~S) continue;
                        if (generic_ptr[g]) continue;                          // (never part of a tile)
                        bool ok = true;
                        for (int c = 0; c < nb && ok; ++c)
                            if (((ubits[g] >> c) & 1) && (h[c] >= queue[c].size() || queue[c][h[c]] != g)) ok = false;
                        if (ok && std::find(ready.begin(), ready.end(), g) == ready.end()) ready.push_back(g);
                    }
                    std::sort(ready.begin(), ready.end());
                    for (int g : ready) {
                        if (count >= max_seg_ops) break;
                        for (int c = 0; c < nb; ++c)
                            if ((ubits[g] >> c) & 1) ++h[c];
                        if (out) out->push_back(g);
                        ++count;
                        progress = true;
                    }
                }
                return count;
            };
            // Order INSIDE a segment: a pass holds R bits in registers and runs every operator whose non-diagonal bits are
            // among them (diagonal operators ride along).  Any order that keeps the per-bit order is the same circuit, so
            // the segment's instructions are listed pass by pass: drain everything that is ready on the current register
            // bits -- several layers deep where the coupling allows -- before the register set grows or a new pass
            // starts.  The pass former below cuts exactly these groups; fewer passes = fewer shared-memory round trips.
            std::vector<uint64_t> ndbits(P.G, 0);              // bits on which the instruction is NOT diagonal
            for (const LogicalOp& l : lops)
                if (!l.diag)
                    for (int b = 0; b < l.nb; ++b) ndbits[l.instr] |= 1ull << l.bits[b];
            auto order_for_passes = [&](std::vector<int>& list) {
                const int n = (int)list.size();
                if (n < 3) return;
                std::vector<int> pos_in(P.G, -1);
                for (int i = 0; i < n; ++i) pos_in[list[i]] = i;
                // predecessors: the previous instruction of the list on each of its bits
                std::vector<std::vector<int>> succ(n);
                std::vector<int> npred(n, 0);
                {
                    std::vector<int> last(nb, -1);
                    for (int i = 0; i < n; ++i) {
                        const int g = list[i];
                        for (int b = 0; b < nb; ++b)
                            if ((ubits[g] >> b) & 1) {
                                if (last[b] >= 0 && (succ[last[b]].empty() || succ[last[b]].back() != i)) { succ[last[b]].push_back(i); ++npred[i]; }
                                last[b] = i;
                            }
                    }
                }
                std::vector<char> emitted(n, 0);
                std::vector<int> out;
                out.reserve(n);
                uint64_t regs = 0;
                auto emit = [&](int i) {
                    emitted[i] = 1; out.push_back(list[i]);
                    for (int j : succ[i]) --npred[j];
                };
                while ((int)out.size() < n) {
                    bool progress = true;
                    while (progress) {
                        progress = false;
                        for (int i = 0; i < n; ++i)               // everything that needs no new register bit
                            if (!emitted[i] && npred[i] == 0 && (ndbits[list[i]] & ~regs) == 0) { emit(i); progress = true; }
                        if (progress) continue;
                        for (int i = 0; i < n; ++i)               // grow the register set with the oldest ready instruction that fits
                            if (!emitted[i] && npred[i] == 0 && __builtin_popcountll(regs | ndbits[list[i]]) <= R) {
                                regs |= ndbits[list[i]]; emit(i); progress = true; break;
                            }
                    }
                    regs = 0;                                     // next pass
                    if ((int)out.size() < n) {
                        bool any = false;
                        for (int i = 0; i < n && !any; ++i)
                            if (!emitted[i] && npred[i] == 0) any = true;
                        if (!any) throw std::runtime_error("internal: pass ordering found no ready instruction");
                        // an instruction with more non-diagonal bits than R (dense operators with reg_bits too small) is
                        // reported by the pass former; emit it here to keep going
                        bool fits = false;
                        for (int i = 0; i < n; ++i)
                            if (!emitted[i] && npred[i] == 0 && __builtin_popcountll(ndbits[list[i]]) <= R) fits = true;
                        if (!fits)
                            for (int i = 0; i < n; ++i)
                                if (!emitted[i] && npred[i] == 0) { emit(i); break; }
                    }
                }
                list.swap(out);
            };
            std::vector<char> done(P.G, 0), generic_instr(P.G, 0);
            for (const LogicalOp& l : lops) if (l.generic) generic_instr[l.instr] = 1;
            int next_undone = 0;
            while ((int)order.size() < P.G) {
                while (next_undone < P.G && done[next_undone]) ++next_undone;
                if (generic_instr[next_undone]) {
                    // a 5..8-bit unitary is a segment of its own; as the oldest pending instruction it is ready
                    for (int c = 0; c < nb; ++c)
                        if ((ubits[next_undone] >> c) & 1) ++head[c];
                    done[next_undone] = 1;
                    order.push_back(next_undone);
                    continue;
                }
                // seed: the oldest pending instruction (it is ready: everything before it on its bits is done)
                uint64_t S = low | ubits[next_undone];
                if (__builtin_popcountll(S) > kfit) throw std::runtime_error("an operation does not fit a tile; increase tile_bits");
                while (__builtin_popcountll(S) < kfit) {
                    int best_bit = -1, best_cnt = -1;
                    long long best_tie = 0;
                    for (int c = 0; c < nb; ++c) {
                        if ((S >> c) & 1) continue;
                        if (head[c] >= queue[c].size()) continue;              // nothing left on this bit
                        std::vector<size_t> h(head);
                        const int cnt = run(S | (1ull << c), h, nullptr);
                        // tie-break: bits coupled to the tile by a pending instruction first, then the oldest pending work
                        const int gp = queue[c][head[c]];
                        const long long tie = ((ubits[gp] & S & ~low) ? (1ll << 40) : 0) - gp;
                        if (cnt > best_cnt || (cnt == best_cnt && tie > best_tie)) { best_cnt = cnt; best_bit = c; best_tie = tie; }
                    }
                    if (best_bit < 0) break;
                    S |= 1ull << best_bit;
                }
                std::vector<int> seg_list;
                const int cnt = run(S, head, &seg_list);
                if (cnt == 0) throw std::runtime_error("internal: the scheduler made no progress");
                if (opt.pass_order) order_for_passes(seg_list);
                for (int g : seg_list) { done[g] = 1; order.push_back(g); }
            }
        }
        // renumber: from here on an instruction's index is its execution position
        P.orig.assign(order.begin(), order.end());
        std::vector<LogicalOp> nl;
        nl.reserve(lops.size());
        std::vector<Instr> ni(P.G);
        P.noisy_instrs.clear();
        for (int pos = 0; pos < P.G; ++pos) {
            const int g = order[pos];
            ni[pos] = P.instrs[g];
            for (int li : ilops[g]) { nl.push_back(lops[li]); nl.back().instr = pos; }
            if (P.draw_pos[g] >= 0) P.noisy_instrs.push_back(pos);
        }
        lops.swap(nl);
        P.instrs.swap(ni);
    }

    // ---- segmentation: maximal runs of consecutive logical ops whose bits fit one tile --------------------------
    struct SegBuild { std::vector<int> tile; size_t lop_begin, lop_end; bool generic = false; };
    std::vector<SegBuild> sb;
    auto fresh_tile = [&]() { std::vector<int> t; for (int b = 0; b < L; ++b) t.push_back(b); return t; };
    auto tile_union = [&](std::vector<int> t, const LogicalOp& l) {
        for (int i = 0; i < l.nb; ++i)
            if (std::find(t.begin(), t.end(), l.bits[i]) == t.end()) t.push_back(l.bits[i]);
        return t;
    };
    {
        // upper bound of what a logical op adds to its segment's shared-memory blob (op records of the op and of a
        // possible group header, its matrices, a share of a pass record)
        auto blob_cost = [&](const LogicalOp& l) {
            const size_t copies = l.bare == l.mat ? 1 : 2;
            const size_t mats = l.diag ? ((size_t)16 << l.nb) * copies + 128
                                       : (l.nb <= 2 ? ((size_t)16 << (2 * l.nb)) * copies : (mode == MODE_DENSITY && l.nb == 4 ? (size_t)1024 * copies : 0));
            return (size_t)(opt.blob_cost_scale * (double)(2 * sizeof(Op) + sizeof(Op) / 2 + mats + sizeof(Pass) / 3));
        };
        size_t cur_cost = sizeof(SegBlob) + sizeof(Pass);
        SegBuild cur{fresh_tile(), 0, 0};
        for (size_t i = 0; i < lops.size(); ++i) {
            if (lops[i].generic) {
                // a 5..8-bit unitary: close the running segment, a segment of its own, start afresh
                if (cur.lop_end > cur.lop_begin) sb.push_back(cur);
                SegBuild gs{fresh_tile(), i, i + 1};
                gs.generic = true;
                sb.push_back(gs);
                cur = SegBuild{fresh_tile(), i + 1, i + 1};
                cur_cost = sizeof(SegBlob) + sizeof(Pass);
                continue;
            }
            const bool new_instr = i == 0 || lops[i].instr != lops[i - 1].instr;
            // an instruction's lops stay in one segment (a jump event reduces over the instruction's qubits right after
            // its last lop): the tile has to take all of them
            std::vector<int> t = tile_union(cur.tile, lops[i]);
            if (new_instr)
                for (size_t j = i + 1; j < lops.size() && lops[j].instr == lops[i].instr; ++j) t = tile_union(t, lops[j]);
            bool force_split = !opt.fuse && new_instr && cur.lop_end > cur.lop_begin;
            if (cur_cost + blob_cost(lops[i]) > (size_t)MAX_BLOB_BYTES - 512 && new_instr && cur.lop_end > cur.lop_begin) force_split = true;
            if ((int)t.size() > kfit || force_split) {
                cur_cost = sizeof(SegBlob) + sizeof(Pass);
                if (cur.lop_end == cur.lop_begin) throw std::runtime_error("an operation does not fit a tile; increase tile_bits");
                sb.push_back(cur);
                cur = SegBuild{fresh_tile(), i, i};
                t = tile_union(cur.tile, lops[i]);
                for (size_t j = i + 1; j < lops.size() && lops[j].instr == lops[i].instr; ++j) t = tile_union(t, lops[j]);
                if ((int)t.size() > kfit) throw std::runtime_error("an operation does not fit a tile; increase tile_bits");
            }
            cur.tile = t; cur.lop_end = i + 1;
            cur_cost += blob_cost(lops[i]);
        }
        if (cur.lop_end > cur.lop_begin || sb.empty()) sb.push_back(cur);
        if (sb.back().generic && mode == MODE_MCWF && !P.norm_preserving) {
            // the squared norm is accumulated by the last sweep of a register-blocked segment: close with an empty one
            SegBuild tail{fresh_tile(), lops.size(), lops.size()};
            sb.push_back(tail);
        }
    }

    // ---- passes and ops -----------------------------------------------------------------------------------------
    for (SegBuild& s : sb) {
        // complete the tile: lowest unused state bits first, then virtual bits
        for (int b = 0; (int)s.tile.size() < kfit && b < P.nbits; ++b)
            if (std::find(s.tile.begin(), s.tile.end(), b) == s.tile.end()) s.tile.push_back(b);
        std::sort(s.tile.begin(), s.tile.end());
        for (int b = P.nbits; (int)s.tile.size() < K; ++b) s.tile.push_back(b);
        Segment seg;
        std::memset(&seg, 0, sizeof(seg));
        int local_of[64];
        for (int i = 0; i < 64; ++i) local_of[i] = -1;
        for (int p = 0; p < K; ++p) { seg.tile_q[p] = (uint8_t)s.tile[p]; local_of[s.tile[p]] = p; }
        {
            std::vector<int> epos(s.tile.begin(), s.tile.begin() + (K - R));
            seg.eruns = make_runs(epos);
            for (int it = 0; it < (1 << R); ++it) {
                unsigned off = 0;
                for (int b = 0; b < R; ++b)
                    if ((it >> b) & 1) off |= 1u << s.tile[K - R + b];
                seg.it_off[it] = off;
                seg.it_swz[it] = (uint16_t)swz(it << (K - R));
            }
            std::vector<int> bpos;
            for (int b = 0; b < P.nbits; ++b)
                if (local_of[b] < 0) bpos.push_back(b);
            seg.bruns = make_runs(bpos);
        }
        seg.pass_begin = (int)P.passes.size();
        seg.instr_lo = s.lop_end > s.lop_begin ? lops[s.lop_begin].instr : (s.lop_begin > 0 ? P.G : 0);
        seg.instr_hi = s.lop_end > s.lop_begin ? lops[s.lop_end - 1].instr + 1 : (s.lop_begin > 0 ? P.G : 0);
        const int seg_id = (int)P.segs.size();
        if (s.generic) {
            const LogicalOp& l = lops[s.lop_begin];
            GenericOp go;
            go.seg = seg_id; go.nb = l.nb;
            for (int i = 0; i < 8; ++i) go.bits[i] = i < l.nb ? l.bits[i] : -1;
            go.mat = pb.add_doubles(l.mat);
            P.generics.push_back(go);
            P.instrs[l.instr].seg = seg_id;
            seg.pad[0] = 1;
            seg.pass_end = seg.pass_begin;
            P.segs.push_back(seg);
            continue;
        }

        std::vector<PassBuild> pbs(1);
        for (size_t i = s.lop_begin; i < s.lop_end; ++i) {
            const LogicalOp& l = lops[i];
            P.instrs[l.instr].seg = seg_id;
            if (!l.diag) {
                if (l.nb > R) throw std::runtime_error("instruction " + std::to_string(l.instr) + ": dense operator on " + std::to_string(l.nb) + " bits needs reg_bits >= " + std::to_string(l.nb));
                std::vector<int> u = pbs.back().regset;
                for (int b = 0; b < l.nb; ++b) {
                    const int p = local_of[l.bits[b]];
                    if (std::find(u.begin(), u.end(), p) == u.end()) u.push_back(p);
                }
                const bool need4 = l.nb >= 3;
                if ((int)u.size() > R) {
                    pbs.emplace_back();
                    u.clear();
                    for (int b = 0; b < l.nb; ++b) u.push_back(local_of[l.bits[b]]);
                }
                pbs.back().regset = u;
                (void)need4;
            }
            pbs.back().lops.push_back(&l);
        }
        for (PassBuild& b : pbs) {
            if (b.lops.empty() && P.G > 0 && pbs.size() > 1) continue;
            Pass ps;
            std::memset(&ps, 0, sizeof(ps));
            // register bits: requested positions, padded with the highest unused tile positions
            std::vector<int> reg = b.regset;
            for (int p = K - 1; (int)reg.size() < R && p >= 0; --p)
                if (std::find(reg.begin(), reg.end(), p) == reg.end()) reg.push_back(p);
            std::sort(reg.begin(), reg.end());
            int reg_of[MAX_TILE_BITS];
            for (int p = 0; p < MAX_TILE_BITS; ++p) reg_of[p] = -1;
            for (int r = 0; r < R; ++r) { ps.regpos[r] = (uint8_t)reg[r]; reg_of[reg[r]] = r; }
            // thread bits: first three with distinct (position mod 3) so a quarter-warp touches 8 bank groups
            std::vector<int> rest;
            for (int p = 0; p < K; ++p)
                if (reg_of[p] < 0) rest.push_back(p);
            std::vector<int> order;
            bool used_res[3] = {false, false, false};
            std::vector<bool> taken(rest.size(), false);
            for (size_t i = 0; i < rest.size() && order.size() < 3; ++i)
                if (!used_res[rest[i] % 3]) { used_res[rest[i] % 3] = true; taken[i] = true; order.push_back(rest[i]); }
            for (size_t i = 0; i < rest.size(); ++i)
                if (!taken[i]) order.push_back(rest[i]);
            for (int i = 0; i < K - R; ++i) ps.tpos[i] = (uint8_t)order[i];
            ps.truns = make_runs(std::vector<int>(order.begin(), order.begin() + (K - R)));
            for (int m = 0; m < (1 << R); ++m) {
                const int off = deposit(m, ps.regpos, R);
                ps.regoff[m] = (uint16_t)off;
                ps.regoff_swz[m] = (uint16_t)swz(off);
            }
            ps.op_begin = (int)P.ops.size();
            ps.instr_lo = b.lops.empty() ? 0 : b.lops.front()->instr;
            ps.instr_hi = b.lops.empty() ? 0 : b.lops.back()->instr + 1;
            auto emit = [&](const LogicalOp& l) -> Op {
                Op op;
                std::memset(&op, 0, sizeof(op));
                op.instr = l.instr;
                int kind = OPK_DIAG;
                uint32_t rb[8] = {0, 0, 0, 0, 0, 0, 0, 0}, diag_rbs = 0;
                const int d = 1 << l.nb;
                if (l.diag) {
                    kind = OPK_DIAG;
                    // table row = hash of the op's thread bits (two multiply-gathers of <= 3 bits), column = its register
                    // bits packed in ascending register order: see diag_mask
                    std::vector<std::pair<int, int>> tb, rg;      // (tile position | register index, op bit)
                    for (int i = 0; i < l.nb; ++i) {
                        const int pos = local_of[l.bits[i]];
                        if (reg_of[pos] < 0) tb.push_back({pos, i}); else rg.push_back({reg_of[pos], i});
                    }
                    std::sort(rg.begin(), rg.end());
                    const int nt = (int)tb.size(), nr = (int)rg.size(), nA = (nt + 1) / 2;
                    if (nt > 6) throw std::runtime_error("diagonal operator with more than 6 thread bits");
                    std::vector<int> posA, posB;
                    for (int i = 0; i < nt; ++i) (i < nA ? posA : posB).push_back(tb[i].first);
                    const Gather gA = find_gather(posA), gB = find_gather(posB);
                    uint32_t regmask = 0;
                    for (auto& r : rg) regmask |= 1u << r.first;
                    kind |= (int)((1u << nA) << 8) | (int)(regmask << 16);
                    diag_rbs = gA.mask | (gB.mask << 16);
                    op.midx_lo = gA.magic; op.midx_hi = gB.magic;
                    auto layout = [&](const std::vector<double>& src) {
                        std::vector<double> dst(src.size());
                        for (int s = 0; s < d; ++s) {
                            uint32_t x = 0, col = 0;
                            for (auto& t : tb) x |= (uint32_t)((s >> t.second) & 1) << t.first;
                            for (int j = 0; j < nr; ++j) col |= (uint32_t)((s >> rg[j].second) & 1) << j;
                            const uint32_t row = (((x & gB.mask) * gB.magic) >> 29) * (1u << nA) + (((x & gA.mask) * gA.magic) >> 29);
                            const size_t e = ((size_t)row << nr) | (col ^ diag_col_key(row, nr));
                            dst[2 * e] = src[2 * (size_t)s]; dst[2 * e + 1] = src[2 * (size_t)s + 1];
                        }
                        return dst;
                    };
                    op.mat = pb.add_doubles(layout(l.mat));
                    op.mat_bare = l.bare == l.mat ? op.mat : pb.add_doubles(layout(l.bare));
                } else if (l.nb <= 2) {
                    for (int i = 0; i < l.nb; ++i) rb[i] = (uint32_t)reg_of[local_of[l.bits[i]]];
                    cmat m = read_complex(l.mat.data(), d * d), mb = read_complex(l.bare.data(), d * d);
                    if (l.nb == 2) {
                        kind = OPK_DENSE2; P.has_dense2 = true;
                        // density-matrix superoperator on (row bit, column bit) that never mixes different r XOR c: two
                        // 2x2 blocks (the R = 4 dense kernels carry the block kernel)
                        auto bd2 = [&](const cmat& a) {
                            for (int ro = 0; ro < 4; ++ro)
                                for (int ri = 0; ri < 4; ++ri)
                                    if ((((ro & 1) ^ (ro >> 1)) != ((ri & 1) ^ (ri >> 1))) && a[ro * 4 + ri] != cd(0, 0)) return false;
                            return true;
                        };
                        if (mode == MODE_DENSITY && R == 4 && opt.fuse && l.bits[1] == l.bits[0] + P.nq && bd2(m) && bd2(mb)) {
                            kind = OPK_DENSE2BD;
                            auto blocks = [](const cmat& a) {
                                cmat b(8);
                                for (int dl = 0; dl < 2; ++dl)
                                    for (int r2 = 0; r2 < 2; ++r2)
                                        for (int r = 0; r < 2; ++r) b[dl * 4 + r2 * 2 + r] = a[(r2 + 2 * (r2 ^ dl)) * 4 + (r + 2 * (r ^ dl))];
                                return b;
                            };
                            m = blocks(m); mb = blocks(mb);
                        }
                    } else {
                        const int km = pattern_1q(m, 4e-16), kb = pattern_1q(mb, 4e-16);
                        kind = km == kb ? km : OPK_DENSE1;
                        snap_pattern(m, kind); snap_pattern(mb, kind);
                    }
                    op.mat = pb.add_doubles(flatten(m));
                    op.mat_bare = l.bare == l.mat ? op.mat : pb.add_doubles(flatten(mb));
                } else {
                    // 3- or 4-bit dense operator: re-index to the pass's four register bits
                    if (R < 4) throw std::runtime_error("dense operators on 3-4 bits need reg_bits = 4");
                    kind = OPK_DENSE4;
                    P.has_dense4 = true;
                    int rbit[4];
                    for (int i = 0; i < l.nb; ++i) rbit[i] = reg_of[local_of[l.bits[i]]];
                    auto remap = [&](const std::vector<double>& src) {
                        const cmat m = read_complex(src.data(), d * d);
                        cmat big(256, cd(0, 0));
                        int spect_mask = 15;
                        for (int i = 0; i < l.nb; ++i) spect_mask &= ~(1 << rbit[i]);
                        for (int ro = 0; ro < 16; ++ro)
                            for (int ri = 0; ri < 16; ++ri) {
                                if ((ro & spect_mask) != (ri & spect_mask)) continue;
                                int mo = 0, mi = 0;
                                for (int i = 0; i < l.nb; ++i) { mo |= ((ro >> rbit[i]) & 1) << i; mi |= ((ri >> rbit[i]) & 1) << i; }
                                big[ro * 16 + ri] = m[mo * d + mi];
                            }
                        return flatten(big);
                    };
                    for (int i = 0; i < 4; ++i) rb[i] = (uint32_t)i;
                    std::vector<double> big = remap(l.mat), big_bare = remap(l.bare);
                    // block structure in register coordinates: m = r | (c << 2) with the row bits on registers 0,1 and the
                    // column bits on 2,3 (true for the density-matrix superoperators: row bits are the lower state bits)
                    auto bd4 = [&](const std::vector<double>& a) {
                        for (int ro = 0; ro < 16; ++ro)
                            for (int ri = 0; ri < 16; ++ri)
                                if (((ro & 3) ^ (ro >> 2)) != ((ri & 3) ^ (ri >> 2)) && (a[2 * (ro * 16 + ri)] != 0.0 || a[2 * (ro * 16 + ri) + 1] != 0.0)) return false;
                        return true;
                    };
                    const bool rows_low = l.nb == 4 && mode == MODE_DENSITY && rbit[0] < 2 && rbit[1] < 2 && rbit[2] >= 2 && rbit[3] >= 2 &&
                                          rbit[2] == rbit[0] + 2 && rbit[3] == rbit[1] + 2;
                    if (rows_low && opt.fuse && bd4(big) && bd4(big_bare)) {
                        kind = OPK_DENSE4BD;
                        auto blocks = [](const std::vector<double>& a) {
                            std::vector<double> b(128);
                            for (int dl = 0; dl < 4; ++dl)
                                for (int r2 = 0; r2 < 4; ++r2)
                                    for (int r = 0; r < 4; ++r) {
                                        const int ro = r2 | ((r2 ^ dl) << 2), ri = r | ((r ^ dl) << 2), e = dl * 16 + r2 * 4 + r;
                                        b[2 * e] = a[2 * (ro * 16 + ri)]; b[2 * e + 1] = a[2 * (ro * 16 + ri) + 1];
                                    }
                            return b;
                        };
                        big = blocks(big); big_bare = blocks(big_bare);
                    }
                    op.mat = pb.add_doubles(big);
                    op.mat_bare = l.bare == l.mat ? op.mat : pb.add_doubles(big_bare);
                }
                op.kind_nb = (uint32_t)(kind & 0xf) | ((uint32_t)l.nb << 4) | ((uint32_t)kind & 0x0fff00u);
                if (R == 3) {   // jump-table case of the lean executor (see ncb_common.h)
                    const int k = kind & 0xf;
                    const uint32_t lsel = k == OPK_DIAG ? ((uint32_t)kind >> 16) & 0x7u : k == OPK_RX1 ? 8u + rb[0] : k == OPK_DENSE1 ? 11u + rb[0] : 15u;
                    op.kind_nb |= lsel << 20;
                }
                op.rbs = diag_rbs;
                if (!l.diag)
                    for (int i = 0; i < 8; ++i) op.rbs |= (rb[i] & 0xfu) << (4 * i);
                op.pad = l.instr;
                return op;
            };
            // Peephole groups: runs of adjacent ops that collapse into one table / matrix.  A group is emitted as a
            // header op (the product, flagged in kind_nb bit 31 with the member count in bits 24..30 and the last
            // member's instruction in `pad`) followed by its members; the executor takes the header when the whole
            // group lies inside the instruction range being run, the members otherwise (jump events are rare).
            struct Group { std::vector<const LogicalOp*> members; LogicalOp merged; };
            std::vector<Group> groups;
            auto try_merge = [&](Group& g, const LogicalOp& l) -> bool {
                if (!opt.fuse || g.members.size() >= 60) return false;
                LogicalOp& a = g.merged;
                if (a.diag && l.diag) {
                    std::vector<int> bits(a.bits, a.bits + a.nb);
                    for (int i = 0; i < l.nb; ++i)
                        if (std::find(bits.begin(), bits.end(), l.bits[i]) == bits.end()) bits.push_back(l.bits[i]);
                    if ((int)bits.size() > max_diag_bits(R)) return false;
                    const int nb = (int)bits.size();
                    std::vector<double> tab(2 << nb);
                    for (int idx = 0; idx < (1 << nb); ++idx) {
                        int ia = 0, il = 0;
                        for (int i = 0; i < a.nb; ++i) ia |= ((idx >> (std::find(bits.begin(), bits.end(), a.bits[i]) - bits.begin())) & 1) << i;
                        for (int i = 0; i < l.nb; ++i) il |= ((idx >> (std::find(bits.begin(), bits.end(), l.bits[i]) - bits.begin())) & 1) << i;
                        const cd v = cd(a.mat[2 * ia], a.mat[2 * ia + 1]) * cd(l.mat[2 * il], l.mat[2 * il + 1]);
                        tab[2 * idx] = v.real(); tab[2 * idx + 1] = v.imag();
                    }
                    a.nb = nb;
                    for (int i = 0; i < 8; ++i) a.bits[i] = i < nb ? bits[i] : -1;
                    a.mat = tab; a.bare = tab;
                    return true;
                }
                if (a.nb == 1 && l.nb == 1 && a.bits[0] == l.bits[0]) {      // same single bit: 2x2 product, l after a
                    auto full = [](const LogicalOp& o) {
                        cmat m(4, cd(0, 0));
                        if (o.diag) { m[0] = cd(o.mat[0], o.mat[1]); m[3] = cd(o.mat[2], o.mat[3]); }
                        else m = read_complex(o.mat.data(), 4);
                        return m;
                    };
                    const cmat prod = matmul(full(l), full(a), 2);
                    LogicalOp r = make_lop(a.instr, {a.bits[0]}, prod, prod);
                    if (r.diag) return false;       // (cannot happen: diag x diag is handled above)
                    a = r;
                    return true;
                }
                return false;
            };
            auto disjoint = [](const LogicalOp& x, const LogicalOp& y) {
                for (int i = 0; i < x.nb; ++i)
                    for (int j = 0; j < y.nb; ++j)
                        if (x.bits[i] == y.bits[j]) return false;
                return true;
            };
            for (const LogicalOp* lp : b.lops) {
                bool placed = false;
                if (mode != MODE_DENSITY && opt.fuse) {
                    if (lp->diag) {
                        // a diagonal operator commutes with diagonal operators and with anything on other bits: walk
                        // back to the nearest diagonal group it can join
                        for (int gi = (int)groups.size() - 1; gi >= 0 && !placed; --gi) {
                            Group& g = groups[gi];
                            if (g.merged.diag && try_merge(g, *lp)) { g.members.push_back(lp); placed = true; break; }
                            if (!(g.merged.diag || disjoint(g.merged, *lp))) break;
                        }
                    } else if (!groups.empty() && try_merge(groups.back(), *lp)) {
                        groups.back().members.push_back(lp);
                        placed = true;
                    }
                }
                if (!placed) groups.push_back(Group{{lp}, *lp});
            }
            std::vector<Op> fast_ops;
            for (const Group& g : groups) {
                if (g.members.size() > 1) {
                    Op h = emit(g.merged);
                    h.instr = g.members.front()->instr;
                    h.pad = g.members.back()->instr;
                    h.mat_bare = h.mat;
                    fast_ops.push_back(h);                 // plain op in the fast list
                    h.kind_nb |= 0x80000000u | ((uint32_t)g.members.size() << 25);
                    P.ops.push_back(h);
                    for (const LogicalOp* lp : g.members) P.ops.push_back(emit(*lp));
                } else {
                    const Op o1 = emit(*g.members[0]);
                    P.ops.push_back(o1);
                    fast_ops.push_back(o1);
                }
            }
            // Fast list only (whole passes without a jump event; the general list above keeps instruction granularity):
            // a generic one-bit product M = D1 . X . D2 hands its phases to the nearest phase tables it can reach through
            // operators it commutes with -- D2 backwards, D1 forwards -- and stays behind as the rx-like X: 8 (4 with its
            // coefficient factored out) instead of 16 (12) FP64 instructions per amplitude pair.  Only when BOTH phases
            // find a table (a new table would cost what the split saves).
            if (mode != MODE_DENSITY && opt.fuse && opt.split_generic) {
                std::vector<LogicalOp> fl;
                for (const Group& g : groups) fl.push_back(g.members.size() > 1 ? g.merged : *g.members[0]);
                std::vector<char> changed(fl.size(), 0);
                auto merge_phase = [&](LogicalOp& a, int bit, const cd ph[2]) -> bool {      // a (diagonal) *= diag(ph) on `bit`
                    std::vector<int> bits(a.bits, a.bits + a.nb);
                    int at = (int)(std::find(bits.begin(), bits.end(), bit) - bits.begin());
                    if (at == (int)bits.size()) {
                        if ((int)bits.size() + 1 > max_diag_bits(R)) return false;
                        // (a wider table is a bigger blob: stay within what the segmenter budgeted, one more bit at most once)
                        if ((int)bits.size() + 1 > 3) return false;
                        bits.push_back(bit);
                        std::vector<double> tab(4 << a.nb);
                        for (int idx = 0; idx < (2 << a.nb); ++idx) { tab[2 * idx] = a.mat[2 * (idx & ((1 << a.nb) - 1))]; tab[2 * idx + 1] = a.mat[2 * (idx & ((1 << a.nb) - 1)) + 1]; }
                        a.mat = tab;
                        a.nb += 1;
                        for (int i = 0; i < 8; ++i) a.bits[i] = i < a.nb ? bits[i] : -1;
                    }
                    for (int idx = 0; idx < (1 << a.nb); ++idx) {
                        const cd v = cd(a.mat[2 * idx], a.mat[2 * idx + 1]) * ph[(idx >> at) & 1];
                        a.mat[2 * idx] = v.real(); a.mat[2 * idx + 1] = v.imag();
                    }
                    a.bare = a.mat;
                    return true;
                };
                auto is_one2 = [](const cd ph[2]) { return std::abs(ph[0] - cd(1, 0)) < 1e-15 && std::abs(ph[1] - cd(1, 0)) < 1e-15; };
                for (size_t i = 0; i < fl.size(); ++i) {
                    if (fl[i].diag || fl[i].nb != 1) continue;
                    const cmat M = read_complex(fl[i].mat.data(), 4);
                    if (pattern_1q(M, 4e-16) == OPK_RX1) continue;
                    cd d1[2], d2[2];
                    cmat X;
                    if (!split_dxd(M, d1, X, d2)) continue;
                    const int bit = fl[i].bits[0];
                    auto touches = [&](const LogicalOp& o) { for (int k = 0; k < o.nb; ++k) if (o.bits[k] == bit) return true; return false; };
                    int jb = -1, jf = -1;
                    for (int j = (int)i - 1; j >= 0 && !is_one2(d2); --j) {
                        if (fl[j].diag) { if (touches(fl[j]) || fl[j].nb < 3) { jb = j; break; } continue; }
                        if (touches(fl[j])) break;
                    }
                    for (size_t j = i + 1; j < fl.size() && !is_one2(d1); ++j) {
                        if (fl[j].diag) { if (touches(fl[j]) || fl[j].nb < 3) { jf = (int)j; break; } continue; }
                        if (touches(fl[j])) break;
                    }
                    if ((!is_one2(d2) && jb < 0) || (!is_one2(d1) && jf < 0)) continue;
                    LogicalOp back = jb >= 0 ? fl[jb] : LogicalOp(), fwd = jf >= 0 ? fl[jf] : LogicalOp();
                    if (jb >= 0 && !merge_phase(back, bit, d2)) continue;
                    if (jf >= 0 && !merge_phase(fwd, bit, d1)) continue;
                    if (jb >= 0) { fl[jb] = back; changed[jb] = 1; }
                    if (jf >= 0) { fl[jf] = fwd; changed[jf] = 1; }
                    fl[i] = make_lop(fl[i].instr, {bit}, X, X);
                    changed[i] = 1;
                }
                for (size_t i = 0; i < fl.size(); ++i)
                    if (changed[i]) {
                        const Op keep = fast_ops[i];
                        Op o = emit(fl[i]);
                        o.instr = keep.instr; o.pad = keep.pad; o.mat_bare = o.mat;
                        fast_ops[i] = o;
                    }
            }
            ps.op_end = (int)P.ops.size();
            ps.fop_begin = (int)P.ops.size();
            P.ops.insert(P.ops.end(), fast_ops.begin(), fast_ops.end());
            ps.fop_end = (int)P.ops.size();
            P.passes.push_back(ps);
        }
        seg.pass_end = (int)P.passes.size();
        P.segs.push_back(seg);
    }
    build_blobs(P);
}

// Builds the shared-memory blob of every segment (see SegBlob).
static void build_blobs(Program& P) {
    P.blobs.clear(); P.blob_off.clear(); P.segconst.clear();
    for (const Segment& sg : P.segs) {
        std::vector<Pass> passes(P.passes.begin() + sg.pass_begin, P.passes.begin() + sg.pass_end);
        std::vector<Op> ops;
        std::vector<double> pool;
        std::vector<std::pair<int, int>> moved;      // (global offset, local offset)
        auto local = [&](int goff, int ndoubles, bool keep_global) {
            if (keep_global) return -(goff + 1);
            for (auto& m : moved)
                if (m.first == goff) return m.second;
            if (pool.size() & 1) pool.push_back(0.0);
            const int off = (int)pool.size();
            pool.insert(pool.end(), P.pool.begin() + goff, P.pool.begin() + goff + ndoubles);
            moved.push_back({goff, off});
            return off;
        };
        for (Pass& ps : passes) {
            const int b = (int)ops.size();
            const int gen_count = ps.op_end - ps.op_begin;
            for (int o = ps.op_begin; o < ps.fop_end; ++o) {      // general list, then the fast list right behind it
                Op op = P.ops[o];
                const int kind = op_kind(op), nb = op_nb(op);
                int nd = 0;
                switch (kind) {
                case OPK_DIAG: nd = 2 << nb; break;
                case OPK_DENSE1: case OPK_RX1: nd = 8; break;
                case OPK_DENSE2: nd = 32; break;
                case OPK_DENSE4: nd = 512; break;
                case OPK_DENSE4BD: nd = 128; break;
                case OPK_DENSE2BD: nd = 16; break;
                }
                const bool big = kind == OPK_DENSE4;
                const int m = op.mat, mb = op.mat_bare;
                op.mat = local(m, nd, big);
                op.mat_bare = mb == m ? op.mat : local(mb, nd, big);
                ops.push_back(op);
            }
            ps.op_begin = b; ps.op_end = b + gen_count;
            ps.fop_begin = ps.op_end; ps.fop_end = (int)ops.size();
        }
        // constant-bank copy of the fast lists (see SegConst)
        {
            SegConst sc;
            std::memset(&sc, 0, sizeof(sc));
            bool fits = (int)passes.size() <= SC_MAX_PASSES;
            int nops = 0, ncoef = 0;
            for (size_t pi = 0; pi < passes.size() && fits; ++pi) {
                const Pass& ps = passes[pi];
                PassHot& ph = sc.pass[pi];
                for (int m = 0; m < (1 << MAX_REG_BITS); ++m) ph.regoff_swz[m] = ps.regoff_swz[m];
                ph.instr_lo = ps.instr_lo; ph.instr_hi = ps.instr_hi;
                ph.fop_begin = nops;
                // Lean programs: the pass's ops are re-ordered among commuting neighbours (diagonals commute with each
                // other and with one-bit ops on register bits outside their mask; one-bit ops on different bits
                // commute) so that the rx-like rotations come in groups, and each group becomes ONE fused op (a single
                // dispatch for up to three rotations).  Only this constant-bank list is rewritten; event pieces keep the
                // program order.
                std::vector<Op> flist(ops.begin() + ps.fop_begin, ops.begin() + ps.fop_end);
                const bool lean = P.R == 3 && !P.has_dense2 && !P.has_dense4 && P.opt.fuse && P.opt.fuse_rotations;
                struct FastItem { Op op; std::vector<Op> rot; };    // rot non-empty: fused rotations (ascending bit order)
                std::vector<FastItem> items;
                if (lean) {
                    const int n = (int)flist.size();
                    auto bitmask = [&](const Op& o) { return op_kind(o) == OPK_DIAG ? op_regmask(o) : 1 << op_rb(o, 0); };
                    auto commute = [&](const Op& x, const Op& y) {
                        const bool dx = op_kind(x) == OPK_DIAG, dy = op_kind(y) == OPK_DIAG;
                        if (dx && dy) return true;
                        return (bitmask(x) & bitmask(y)) == 0;
                    };
                    std::vector<char> done(n, 0);
                    auto ready = [&](int j) {
                        for (int i = 0; i < j; ++i)
                            if (!done[i] && !commute(flist[i], flist[j])) return false;
                        return true;
                    };
                    int left = n;
                    while (left > 0) {
                        bool any = false;
                        for (int j = 0; j < n; ++j)                  // every ready diagonal first
                            if (!done[j] && op_kind(flist[j]) == OPK_DIAG && ready(j)) { items.push_back({flist[j], {}}); done[j] = 1; --left; any = true; }
                        std::vector<int> grp;                        // then the ready rx-like rotations, as one op
                        int used = 0;
                        for (int j = 0; j < n; ++j)
                            if (!done[j] && op_kind(flist[j]) == OPK_RX1 && ready(j) && !(used & bitmask(flist[j]))) { grp.push_back(j); used |= bitmask(flist[j]); }
                        if (grp.size() >= 2) {
                            FastItem it{flist[grp[0]], {}};
                            std::sort(grp.begin(), grp.end(), [&](int x, int y) { return op_rb(flist[x], 0) < op_rb(flist[y], 0); });
                            for (int j : grp) { it.rot.push_back(flist[j]); done[j] = 1; --left; }
                            items.push_back(it);
                            any = true;
                        } else {
                            for (int j = 0; j < n; ++j)              // a single rotation / generic one-bit op: as it is
                                if (!done[j] && op_kind(flist[j]) != OPK_DIAG && ready(j)) { items.push_back({flist[j], {}}); done[j] = 1; --left; any = true; break; }
                        }
                        if (!any) throw std::runtime_error("internal: fast-list reordering made no progress");
                    }
                } else {
                    for (const Op& o : flist) items.push_back({o, {}});
                }
                for (size_t fi = 0; fi < items.size() && fits; ++fi) {
                    if (nops >= SC_MAX_OPS) { fits = false; break; }
                    Op op = items[fi].op;
                    if (!items[fi].rot.empty()) {
                        const std::vector<Op>& rot = items[fi].rot;
                        const int need = 4 * (int)rot.size();
                        if (ncoef + need > SC_MAX_COEF) { fits = false; break; }
                        int bits = 0;
                        for (size_t r = 0; r < rot.size(); ++r) {
                            const double* m = pool.data() + rot[r].mat;
                            sc.coef[ncoef + 4 * r] = m[0]; sc.coef[ncoef + 4 * r + 1] = m[3]; sc.coef[ncoef + 4 * r + 2] = m[5]; sc.coef[ncoef + 4 * r + 3] = m[6];
                            bits |= 1 << op_rb(rot[r], 0);
                        }
                        const uint32_t lsel = rot.size() == 3 ? (uint32_t)LSEL_X3 : (uint32_t)(LSEL_X2 + (bits == 3 ? 0 : bits == 5 ? 1 : 2));
                        op.kind_nb = (op.kind_nb & ~(0x1fu << 20)) | (lsel << 20);
                        op.mat = ncoef; op.mat_bare = ncoef;
                        ncoef += need;
                        sc.ops[nops++] = op;
                        continue;
                    }
                    const int kind = op_kind(op);
                    if (kind == OPK_RX1 || kind == OPK_DENSE1) {
                        const int need = kind == OPK_RX1 ? 4 : 8;
                        if (ncoef + need > SC_MAX_COEF) { fits = false; break; }
                        const double* m = pool.data() + op.mat;       // 4 complex, row-major
                        if (kind == OPK_RX1) { sc.coef[ncoef] = m[0]; sc.coef[ncoef + 1] = m[3]; sc.coef[ncoef + 2] = m[5]; sc.coef[ncoef + 3] = m[6]; }
                        else for (int i = 0; i < 8; ++i) sc.coef[ncoef + i] = m[i];
                        op.mat = ncoef; op.mat_bare = ncoef;
                        ncoef += need;
                    }
                    sc.ops[nops++] = op;
                }
                ph.fop_end = nops;
            }
            sc.valid = fits ? 1 : 0;
            sc.npass = (int)passes.size(); sc.nops = nops; sc.ncoef = ncoef;
            // direct store of the last pass
            if (fits && !passes.empty() && P.nbits > P.K && P.opt.direct && (P.R == 3 || P.R == 4) && P.mode != MODE_DENSITY && !P.has_dense2 && !P.has_dense4) {
                const int Lb = std::min(std::max(P.opt.low_bits, 0), std::min(P.K - 4, P.nbits));
                auto describe_pass = [&](const Pass& ps, uint32_t* goff, BitRuns& gruns) -> bool {
                    bool ok = Lb >= 2;
                    // every 32-byte sector a warp's store touches must be completed by the same warp: state bits 0 and 1
                    // have to be lane bits (in order) or register bits (the thread writes the neighbours itself)
                    for (int i = 0, lane = 0; i < Lb; ++i) {
                        if (sg.tile_q[i] != i) ok = false;
                        bool is_reg = false;
                        for (int r = 0; r < P.R; ++r) is_reg = is_reg || ps.regpos[r] == i;
                        if (is_reg) continue;
                        if (ps.tpos[lane] != i) ok = false;
                        ++lane;
                    }
                    for (int m = 0; m < (1 << P.R); ++m) {
                        uint32_t off = 0;
                        for (int r = 0; r < P.R; ++r)
                            if ((m >> r) & 1) off |= 1u << sg.tile_q[ps.regpos[r]];
                        goff[m] = off;
                    }
                    std::vector<int> gpos;
                    for (int i = 0; i < P.K - P.R; ++i) gpos.push_back(sg.tile_q[ps.tpos[i]]);
                    gruns = make_runs(gpos);
                    return ok;
                };
                sc.direct = describe_pass(passes.back(), sc.goff_last, sc.gruns_last) ? 1 : 0;
                sc.direct_first = describe_pass(passes.front(), sc.goff_first, sc.gruns_first) ? 1 : 0;
            }
            P.segconst.push_back(sc);
        }
        SegBlob h;
        std::memset(&h, 0, sizeof(h));
        h.seg = sg;
        h.seg.pass_begin = 0; h.seg.pass_end = (int)passes.size();
        h.npass = (int)passes.size(); h.nops = (int)ops.size(); h.pool_doubles = (int)pool.size();
        auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
        h.pass_off = (int)align16(sizeof(SegBlob));
        h.op_off = (int)align16(h.pass_off + passes.size() * sizeof(Pass));
        h.pool_off = (int)align16(h.op_off + ops.size() * sizeof(Op));
        h.bytes = (int)align16(h.pool_off + pool.size() * sizeof(double));
        if (h.bytes > MAX_BLOB_BYTES) throw BlobOverflow("internal: segment program of " + std::to_string(h.bytes) + " bytes exceeds the shared-memory staging area");
        const size_t at = P.blobs.size();
        P.blob_off.push_back((int64_t)at);
        P.blobs.resize(at + h.bytes, 0);
        std::memcpy(P.blobs.data() + at, &h, sizeof(h));
        if (!passes.empty()) std::memcpy(P.blobs.data() + at + h.pass_off, passes.data(), passes.size() * sizeof(Pass));
        if (!ops.empty()) std::memcpy(P.blobs.data() + at + h.op_off, ops.data(), ops.size() * sizeof(Op));
        if (!pool.empty()) std::memcpy(P.blobs.data() + at + h.pool_off, pool.data(), pool.size() * sizeof(double));
    }
}

std::string describe_program(const Program& P) {
    std::ostringstream o;
    o << "mode=" << P.mode << " nq=" << P.nq << " nbits=" << P.nbits << " K=" << P.K << " R=" << P.R << " G=" << P.G
      << " segments=" << P.segs.size() << " passes=" << P.passes.size() << " ops=" << P.ops.size()
      << " channels=" << P.chans.size() << " noisy=" << P.noisy_instrs.size() << " draws=" << P.num_draws
      << " norm_preserving=" << P.norm_preserving << " dense2=" << P.has_dense2 << " dense4=" << P.has_dense4
      << " pool=" << P.pool.size() << " blobs=" << P.blobs.size() << " reorder=" << P.opt.reorder << " attempts=" << P.compile_attempts;
    {
        static const char* names[] = {"diag", "dense1", "dense2", "dense4", "rx1", "dense4bd", "dense2bd", "?"};
        int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, groups = 0;
        for (const Op& op : P.ops) { ++cnt[op_kind(op) & 7]; groups += (int)(op.kind_nb >> 31); }
        o << " kinds{";
        for (int k = 0; k < 8; ++k) if (cnt[k]) o << names[k] << ":" << cnt[k] << " ";
        o << "groups:" << groups << "}";
        int fast = 0, fdiag = 0;
        for (const Pass& ps : P.passes)
            for (int i = ps.fop_begin; i < ps.fop_end; ++i) { ++fast; fdiag += op_kind(P.ops[i]) == OPK_DIAG; }
        int valid = 0, direct = 0;
        for (const SegConst& sc : P.segconst) { valid += sc.valid; direct += sc.direct; }
        o << " fast_ops=" << fast << " (diag " << fdiag << ") const_segments=" << valid << "/" << P.segconst.size() << " direct_segments=" << direct;
    }
    o << "\n";
    for (size_t s = 0; s < P.segs.size(); ++s) {
        const Segment& g = P.segs[s];
        o << "seg " << s << " instr[" << g.instr_lo << "," << g.instr_hi << ") tile{";
        for (int p = 0; p < P.K; ++p) o << (p ? "," : "") << (int)g.tile_q[p];
        o << "} passes " << g.pass_end - g.pass_begin;
        // fast op list of every pass: d<regmask> diagonal, x<bit> rx-like, u<bit> generic one-bit, D dense 4x4 / 16x16
        if (s < P.segconst.size() && P.segconst[s].valid) {
            const SegConst& sc = P.segconst[s];
            o << " ops";
            for (int p = 0; p < sc.npass; ++p) {
                o << " [";
                for (int i = sc.pass[p].fop_begin; i < sc.pass[p].fop_end; ++i) {
                    const Op& op = sc.ops[i];
                    const int k = op_kind(op);
                    if (k == OPK_DIAG) o << "d" << op_regmask(op);
                    else if (k == OPK_RX1) o << "x" << op_rb(op, 0);
                    else if (k == OPK_DENSE1) o << "u" << op_rb(op, 0);
                    else o << "D";
                    if (i + 1 < sc.pass[p].fop_end) o << " ";
                }
                o << "]";
            }
        }
        o << "\n";
    }
    return o.str();
}

}  // namespace ncb
