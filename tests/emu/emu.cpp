// TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// CPU emulation of the tile kernel (noisycircuits_b200/csrc/ncb_kernels.cuh): the same host compiler
// (ncb_compile.cpp), the same batch scheduler (ncb_schedule.h) and the same per-thread executors (ncb_common.h) are
// driven thread by thread, CTA by CTA, launch by launch.  It exists so that the host planner, the scheduler, the tile
// geometry and the register-blocked index maps can be tested on machines without a GPU (pytest -m "not gpu").
// CUDA-only pieces (cp.async, warp shuffles, launch configuration) are not covered here; the -m gpu tests are.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ncb200.h"
#include "../../noisycircuits_b200/csrc/ncb_compile.h"
#include "../../noisycircuits_b200/csrc/ncb_schedule.h"

using namespace ncb;

namespace {

thread_local std::string g_err;

struct Emu {
    Program P;
};

template <int R>
void emulate_work(const Program& P, const Work& w, int seg, std::vector<cplx>& states, std::vector<double>& red_part,
                  std::vector<double>& norm_part, const std::vector<Decision>& decisions) {
    const int K = P.K, nthreads = 1 << (K - R);
    const int tile_shift = P.nbits > K ? P.nbits - K : 0, ntiles = 1 << tile_shift;
    // the staged copy of the segment program (what the tile kernel reads from shared memory)
    const unsigned char* braw = P.blobs.data() + P.blob_off[seg];
    const SegBlob& blob = *reinterpret_cast<const SegBlob*>(braw);
    const Segment& sg = blob.seg;
    const Pass* passes = reinterpret_cast<const Pass*>(braw + blob.pass_off);
    const Op* ops = reinterpret_cast<const Op*>(braw + blob.op_off);
    const double* pool = reinterpret_cast<const double*>(braw + blob.pool_off);
    if (blob.bytes > MAX_BLOB_BYTES) throw std::runtime_error("blob too large");
    const long long limit = 1ll << P.nbits;
    std::vector<cplx> sm(1 << K);
    for (int o = 0; o < ntiles; ++o) {
        const long long base = apply_runs(sg.bruns, (unsigned)o);
        cplx* st = states.data() + ((long long)w.slot << P.nbits) + base;
        const cplx* st_src = states.data() + ((long long)w.src << P.nbits) + base;
        // load
        for (int tid = 0; tid < nthreads; ++tid) {
            const unsigned offA = apply_runs(sg.eruns, (unsigned)tid);
            const int sA = swz(tid);
            for (int it = 0; it < (1 << R); ++it) {
                const unsigned off = offA | sg.it_off[it];
                cplx v{0.0, 0.0};
                if (!(w.flags & WF_INIT) && off < limit) v = st_src[off];
                sm[sA ^ sg.it_swz[it]] = v;
            }
        }
        if ((w.flags & WF_INIT) && o == 0) sm[swz(0)] = cplx{1.0, 0.0};
        // jump operator
        auto apply_event = [&](int instr, int k, double scale) {
            const Instr& in = P.instrs[instr];
            const Channel& ch = P.chans[in.channel];
            const cplx* M = reinterpret_cast<const cplx*>(P.pool.data() + ch.kraus) + k * ch.dim * ch.dim;
            const int p0 = tile_pos(sg, K, in.q0), p1 = in.q1 >= 0 ? tile_pos(sg, K, in.q1) : -1;
            for (int tid = 0; tid < nthreads; ++tid) tile_apply_dense(sm.data(), K, tid, nthreads, ch.dim, p0, p1, M, scale);
        };
        if (w.flags & WF_PRO_STATIC) apply_event(w.ev_instr, w.ev_k, 1.0);
        else if (w.flags & WF_PRO_DYNAMIC) apply_event(w.ev_instr, decisions[w.slot].k, decisions[w.slot].scale);
        // passes (same control flow as run_range in ncb_kernels.cuh); WF_MID: a jump applied on the way
        int lo_i = w.instr_lo;
        for (int part = (w.flags & WF_MID) ? 0 : 1; part < 2; ++part) {
            const int hi_i = part == 0 ? w.mid_instr + 1 : w.instr_hi;
            const bool bare_last = part == 0 ? true : (w.flags & WF_BARE_LAST) != 0;
            for (int p = 0; p < blob.npass; ++p) {
                const Pass& ps = passes[p];
                if (ps.instr_lo >= hi_i) break;
                if (ps.instr_hi > lo_i && ps.instr_hi > ps.instr_lo)
                    for (int tid = 0; tid < nthreads; ++tid) {
                        // the lean kernels (FEAT = 0) run the programs without dense two-qubit operators, like the engine
                        const int jb = (int)apply_runs(ps.truns, (unsigned)tid);
                        if (R == 3 && !P.has_dense2 && !P.has_dense4)
                            run_pass_thread<R, 0, true>(sm.data(), K, jb, ps, ops, pool, P.pool.data(), lo_i, hi_i, bare_last, P.segconst[seg], p);
                        else
                            run_pass_thread<R, 3, true>(sm.data(), K, jb, ps, ops, pool, P.pool.data(), lo_i, hi_i, bare_last, P.segconst[seg], p);
                    }
                if (ps.instr_hi > hi_i) break;
            }
            if (part == 0) {
                apply_event(w.mid_instr, (int)((unsigned)w.flags >> 16), 1.0);
                lo_i = w.mid_instr + 1;
            }
        }
        // reductions
        if (w.flags & WF_REDUCE) {
            const Instr& in = P.instrs[w.instr_hi - 1];
            const int p0 = tile_pos(sg, K, in.q0), p1 = in.q1 >= 0 ? tile_pos(sg, K, in.q1) : -1;
            double acc[32];
            for (double& a : acc) a = 0.0;
            for (int tid = 0; tid < nthreads; ++tid) {
                if (p1 < 0) reduce_rho_thread<2>(sm.data(), K, tid, nthreads, p0, -1, acc);
                else reduce_rho_thread<4>(sm.data(), K, tid, nthreads, p0, p1, acc);
            }
            double* dst = red_part.data() + ((long long)w.slot * ntiles + o) * 32;
            for (int i = 0; i < 32; ++i) dst[i] = acc[i];
        }
        if (w.flags & WF_NORM) {
            double s = 0.0;
            for (int i = 0; i < (1 << K); ++i) s += cnorm2(sm[i]);
            norm_part[(long long)w.slot * ntiles + o] = s;
        }
        // store
        for (int tid = 0; tid < nthreads; ++tid) {
            const unsigned offA = apply_runs(sg.eruns, (unsigned)tid);
            const int sA = swz(tid);
            for (int it = 0; it < (1 << R); ++it) {
                const unsigned off = offA | sg.it_off[it];
                if (off < limit) st[off] = sm[sA ^ sg.it_swz[it]];
            }
        }
    }
}

// the direct kernel (ncb_kernels.cuh: direct_kernel): tile loaded like the tile kernel, last pass stored from registers
void emulate_direct_work(const Program& P, const Work& w, int seg, std::vector<cplx>& states) {
    constexpr int R = 3;
    if (P.R != 3) throw std::runtime_error("direct launch on a program with reg_bits != 3");
    const int K = P.K, nthreads = 1 << (K - R);
    const int tile_shift = P.nbits - K, ntiles = 1 << tile_shift;
    const unsigned char* braw = P.blobs.data() + P.blob_off[seg];
    const SegBlob& blob = *reinterpret_cast<const SegBlob*>(braw);
    const Segment& sg = blob.seg;
    const Pass* passes = reinterpret_cast<const Pass*>(braw + blob.pass_off);
    const Op* ops = reinterpret_cast<const Op*>(braw + blob.op_off);
    const double* pool = reinterpret_cast<const double*>(braw + blob.pool_off);
    const SegConst& sc = P.segconst[seg];
    if (!sc.valid || !sc.direct) throw std::runtime_error("direct launch on a segment that does not allow it");
    if (w.flags != 0 || w.instr_lo > sg.instr_lo || w.instr_hi < sg.instr_hi) throw std::runtime_error("direct launch with a partial work");
    std::vector<cplx> sm((size_t)1 << K);
    for (int o = 0; o < ntiles; ++o) {
        const long long base = apply_runs(sg.bruns, (unsigned)o);
        const cplx* src = states.data() + ((long long)w.src << P.nbits) + base;
        cplx* dst = states.data() + ((long long)w.slot << P.nbits) + base;
        for (int tid = 0; tid < nthreads; ++tid) {
            const unsigned offA = apply_runs(sg.eruns, (unsigned)tid);
            const int sA = swz(tid);
            for (int it = 0; it < (1 << R); ++it) sm[sA ^ sg.it_swz[it]] = src[offA | sg.it_off[it]];
        }
        for (int p = 0; p < sc.npass; ++p) {
            const bool last = p == sc.npass - 1;
            // the last pass reads the tile into "registers" for every thread before any of them stores: the kernel's
            // barrier in after_load; here the stores go to global memory, which the loads never read (src was consumed)
            for (int tid = 0; tid < nthreads; ++tid)
                run_pass_thread<R, 0, true>(sm.data(), K, (int)apply_runs(passes[p].truns, (unsigned)tid), passes[p], ops, pool, P.pool.data(),
                                            sg.instr_lo, sg.instr_hi, false, sc, p, last ? dst : nullptr,
                                            last ? apply_runs(sc.gruns_last, (unsigned)tid) : 0u);
        }
    }
}

// a segment that is one 5..8-bit unitary (generic_unitary_kernel): gather, mat-vec, scatter per group
void emulate_generic(const Program& P, const Work& w, int seg, std::vector<cplx>& states) {
    const GenericOp* go = nullptr;
    for (const GenericOp& g : P.generics) if (g.seg == seg) go = &g;
    if (!go) throw std::runtime_error("generic launch on an ordinary segment");
    const int k = go->nb, dimk = 1 << k;
    const cplx* U = reinterpret_cast<const cplx*>(P.pool.data() + go->mat);
    unsigned tmask = 0;
    for (int i = 0; i < k; ++i) tmask |= 1u << go->bits[i];
    std::vector<cplx> in(dimk), out(dimk);
    for (long long g = 0; g < (1ll << (P.nbits - k)); ++g) {
        long long gg = g;
        unsigned base = 0;
        for (int b = 0; b < P.nbits; ++b)
            if (!((tmask >> b) & 1u)) { base |= (unsigned)(gg & 1) << b; gg >>= 1; }
        std::vector<long long> off(dimk);
        for (int r = 0; r < dimk; ++r) {
            unsigned sp = 0;
            for (int i = 0; i < k; ++i) sp |= (unsigned)((r >> i) & 1) << go->bits[i];
            off[r] = (long long)(base | sp);
            in[r] = (w.flags & WF_INIT) ? cplx{off[r] == 0 ? 1.0 : 0.0, 0.0} : states[((size_t)w.src << P.nbits) + off[r]];
        }
        for (int r = 0; r < dimk; ++r) {
            cplx acc = cmul(U[(size_t)r * dimk], in[0]);
            for (int c = 1; c < dimk; ++c) acc = cfma(U[(size_t)r * dimk + c], in[c], acc);
            out[r] = acc;
        }
        for (int r = 0; r < dimk; ++r) states[((size_t)w.slot << P.nbits) + off[r]] = out[r];
    }
}

void run_plan(const Program& P, const BatchPlan& plan, int B, std::vector<cplx>& states, std::vector<double>& norm_part) {
    const int tile_shift = P.nbits > P.K ? P.nbits - P.K : 0, ntiles = 1 << tile_shift;
    std::vector<double> red_part((size_t)B * ntiles * 32, 0.0);
    norm_part.assign((size_t)B * ntiles, 0.0);
    std::vector<Decision> decisions(B);
    for (const LaunchStep& st : plan.steps) {
        if (st.kind == 2) {
            for (int i = 0; i < st.count; ++i) emulate_generic(P, plan.works[st.begin + i], st.seg, states);
        } else if (st.kind == 0) {
            for (int i = 0; i < st.count; ++i) {
                const Work& w = plan.works[st.begin + i];
                if (st.direct && P.R == 3) { emulate_direct_work(P, w, st.seg, states); continue; }      // (as Engine::launch_step)
                if (P.R == 4) emulate_work<4>(P, w, st.seg, states, red_part, norm_part, decisions);
                else emulate_work<3>(P, w, st.seg, states, red_part, norm_part, decisions);
            }
        } else {
            for (int i = 0; i < st.count; ++i) {
                const DecideItem& d = plan.decides[st.begin + i];
                double rho[32];
                for (int e = 0; e < 32; ++e) {
                    double s = 0.0;
                    for (int o = 0; o < ntiles; ++o) s += red_part[((size_t)d.slot * ntiles + o) * 32 + e];
                    rho[e] = s;
                }
                const Channel& ch = P.chans[P.instrs[d.instr].channel];
                mirror_rho(rho, ch.dim);
                double scale;
                decisions[d.slot].k = decide_branch(rho, ch, P.pool.data(), d.u, &scale);
                decisions[d.slot].scale = scale;
            }
        }
    }
}

CircuitView view_of(const ncb_circuit* c) {
    CircuitView v;
    v.num_qubits = c->num_qubits; v.num_instructions = c->num_instructions;
    v.opcode = c->opcode; v.q0 = c->q0; v.q1 = c->q1; v.theta = c->theta; v.channel = c->channel;
    v.unitary_offset = c->unitary_offset; v.unitary_nq = c->unitary_nq; v.unitary_qubits = c->unitary_qubits;
    v.unitary_pool = c->unitary_pool; v.num_channels = c->num_channels; v.channel_offset = c->channel_offset;
    v.channel_count = c->channel_count; v.channel_dim = c->channel_dim; v.kraus_pool = c->kraus_pool;
    return v;
}

}  // namespace

extern "C" {

const char* emu_last_error() { return g_err.c_str(); }

void* emu_create(const ncb_circuit* c, int mode, int tile_bits, int reg_bits, int low_bits, int fuse, int reorder) {
    try {
        Emu* e = new Emu();
        CompileOptions opt;
        opt.tile_bits = tile_bits; opt.resident_bits = tile_bits; opt.reg_bits = reg_bits; opt.low_bits = low_bits; opt.fuse = fuse; opt.reorder = reorder;
        if (const char* e = getenv("NCB_SPLIT_U")) opt.split_generic = atoi(e);      // (the engine reads the same switch)
        compile_program(view_of(c), mode, opt, e->P);
        return e;
    } catch (const std::exception& ex) { g_err = ex.what(); return nullptr; }
}
void emu_destroy(void* h) { delete static_cast<Emu*>(h); }
void emu_order(void* h, int* out) {
    const Program& P = static_cast<Emu*>(h)->P;
    for (int i = 0; i < P.G; ++i) out[i] = P.orig[i];
}

long long emu_query(void* h, int what) {
    const Program& P = static_cast<Emu*>(h)->P;
    switch (what) {
    case 0: return P.K; case 1: return P.R; case 2: return (long long)P.segs.size(); case 3: return (long long)P.passes.size();
    case 4: return (long long)P.ops.size(); case 5: return P.norm_preserving; case 6: return P.nbits; case 7: return P.has_dense4;
    }
    return -1;
}

long long emu_describe(void* h, char* buf, long long len) {
    const std::string s = describe_program(static_cast<Emu*>(h)->P);
    if (buf && len > 0) { const size_t n = std::min<size_t>(s.size(), (size_t)len - 1); std::memcpy(buf, s.data(), n); buf[n] = 0; }
    return (long long)s.size() + 1;
}

// Runs `count` trajectories (explicit uniforms [count][G]) as ONE batch and returns their final normalised states
// [count][2^n] (complex interleaved), the number of events / state-dependent events, and the sweeps issued.
int emu_mcwf_states(void* h, int count, const double* uniforms, double* states_out, long long* stats, int share) {
    try {
        const Program& P = static_cast<Emu*>(h)->P;
        if (P.mode != MODE_MCWF) throw std::runtime_error("not an MCWF program");
        const long long dim = 1ll << P.nbits;
        std::vector<std::vector<Event>> events(count);
        long long nev = 0, nhard = 0;
        for (int b = 0; b < count; ++b) {
            plan_events(P, [&](int i) { return uniforms[(size_t)b * P.G + i]; }, events[b]);
            nev += (long long)events[b].size();
            for (const Event& e : events[b]) nhard += e.kmin != e.kmax;
        }
        BatchPlan plan;
        schedule_batch(P, events, true, !P.norm_preserving, plan, share != 0);
        // unused slots start as NaN: a schedule that reads a slot nobody wrote cannot pass
        std::vector<cplx> states((size_t)plan.nslots * dim, cplx{std::nan(""), std::nan("")});
        std::vector<double> norm_part;
        run_plan(P, plan, plan.nslots, states, norm_part);
        const int ntiles = 1 << (P.nbits > P.K ? P.nbits - P.K : 0);
        for (int b = 0; b < count; ++b) {
            // a trajectory without any jump ends in the shared trunk's state
            const int sl = plan.weight[b] != 0.0 ? b : plan.trunk_slot;
            if (sl < 0) throw std::runtime_error("schedule lost a trajectory");
            const cplx* fin = states.data() + (size_t)sl * dim;
            double n2 = 0.0;
            for (long long j = 0; j < dim; ++j) n2 += cnorm2(fin[j]);
            if (!P.norm_preserving) {     // the WF_NORM partials must agree with the direct norm
                double s = 0.0;
                for (int o = 0; o < ntiles; ++o) s += norm_part[(size_t)sl * ntiles + o];
                if (std::fabs(s - n2) > 1e-12 * std::max(1.0, n2)) throw std::runtime_error("WF_NORM partials disagree with the state norm");
            }
            const double f = 1.0 / std::sqrt(n2);
            const double gx = P.global_phase[0] * f, gy = P.global_phase[1] * f;
            for (long long j = 0; j < dim; ++j) {
                const cplx a = fin[j];
                states_out[2 * ((size_t)b * dim + j)] = a.x * gx - a.y * gy;
                states_out[2 * ((size_t)b * dim + j) + 1] = a.x * gy + a.y * gx;
            }
        }
        if (stats) { stats[0] = nev; stats[1] = nhard; stats[2] = plan.sweeps; stats[3] = (long long)plan.steps.size(); }
        return 0;
    } catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}

// Deterministic programs (PURE / DENSITY): final state vector of 2^nbits amplitudes.
int emu_run_deterministic(void* h, double* state_out) {
    try {
        const Program& P = static_cast<Emu*>(h)->P;
        const long long dim = 1ll << P.nbits;
        std::vector<std::vector<Event>> events(1);
        BatchPlan plan;
        schedule_batch(P, events, true, false, plan);
        std::vector<cplx> states(dim, cplx{0.0, 0.0});
        std::vector<double> norm_part;
        run_plan(P, plan, 1, states, norm_part);
        for (long long j = 0; j < dim; ++j) {
            state_out[2 * j] = states[j].x * P.global_phase[0] - states[j].y * P.global_phase[1];
            state_out[2 * j + 1] = states[j].x * P.global_phase[1] + states[j].y * P.global_phase[0];
        }
        return 0;
    } catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}

// host-side Philox, for cross-checking the device draws
double emu_philox_uniform(unsigned long long seed, unsigned long long traj, unsigned instr) { return philox_uniform(seed, traj, instr); }

}  // extern "C"
