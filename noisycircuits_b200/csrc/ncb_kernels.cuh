// sm_100a kernels of the B200 noisy-circuit engine.
//
//   tile_kernel      one CTA = one 2^K-amplitude tile of one trajectory, staged in shared memory: load (cp.async,
//                    128-bit, runs of >= 2^low_bits amplitudes), optional jump operator, the segment's register-blocked
//                    passes, optional reduced-density-matrix / norm reduction, store.  One launch sweeps the state of
//                    every trajectory of the batch once (one HBM read + one HBM write per amplitude), whatever the
//                    number of fused instructions.  Handles the works that carry a jump event (pieces of a segment).
//   direct_kernel    the same sweep for the works that run a whole segment without an event (most of them): lean
//                    operator set only, last pass stored from registers to HBM, next tile's load overlapped with it.
//   resident_kernel  states of <= 2^K amplitudes: the whole trajectory lives in one CTA's shared memory; persistent
//                    CTAs loop over trajectories, plan their jump events with Philox draws, handle them inline and
//                    accumulate |psi|^2 on chip.  HBM is touched only for the final probabilities.
//   decide_kernel    branch decision of the state-dependent jump events from the reduced density matrices.
//   accumulate / norm / diag / measurement-error kernels: the tail of the path.
#pragma once
#if !defined(__CUDACC_RTC__)
#include <cuda_runtime.h>
#endif

#include "ncb_common.h"

namespace ncb {

struct DevProgram {
    const Op* ops;
    const Pass* passes;
    const Segment* segs;
    const Instr* instrs;
    const Channel* chans;
    const double* pool;
    const int32_t* noisy;   // execution positions with a channel of >= 2 operators
    const int32_t* orig;    // orig[pos]: the caller's instruction index (keys the Philox draw)
    int n_noisy, G, nbits, K, R, nseg, npass;
};

constexpr int RED_VALUES = 32;   // doubles per reduced-density-matrix partial
constexpr int JB_TABLE_PASSES = 8;   // passes per segment whose per-thread base index is tabulated in shared memory

// ---------------------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block reduction of NV per-thread values: fixed shuffle tree inside a warp, warps summed in order.
// s_buf: [32 warps][NV] doubles of shared memory.  Result valid in out[0..NV) on threads < NV after the call.
template <int NV>
__device__ __forceinline__ void block_reduce(const double* vals, double* s_buf, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const double s = warp_sum(vals[v]);
        if (lane == 0) s_buf[warp * NV + v] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w) s += s_buf[w * NV + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// Runs the ops with instruction index in [lo, hi) of the passes [*pcur, pend) on the tile in shared memory.
// *pcur is advanced past the passes that are entirely below hi.  passes/ops/pool may live in shared memory (the
// tile kernel's staged segment blob) or in global memory (the resident kernel).
// jb_tab (optional, shared memory): jb_tab[p * blockDim.x + tid] = this thread's base tile index in pass p for
// p < jb_cap (computed once per CTA: it does not depend on the tile).
template <int R, int FEAT, bool USE_SC>
__device__ __forceinline__ void run_range(cplx* sm, int K, const Pass* passes, const Op* ops, const double* pool,
                                          const double* pool_global, int* pcur, int pend, int lo, int hi, bool bare_last,
                                          const SegConst& sc, const unsigned short* jb_tab = nullptr, int jb_cap = 0) {
    int p = *pcur;
    for (; p < pend; ++p) {
        const Pass* ps = passes + p;
        const bool cst = USE_SC && sc.valid != 0;
        const int plo = cst ? sc.pass[p].instr_lo : ps->instr_lo, phi = cst ? sc.pass[p].instr_hi : ps->instr_hi;
        if (plo >= hi) break;
        if (phi > lo && phi > plo) {
            const int jb = p < jb_cap ? (int)jb_tab[p * blockDim.x + threadIdx.x] : (int)apply_runs(ps->truns, threadIdx.x);
            run_pass_thread<R, FEAT, USE_SC>(sm, K, jb, *ps, ops, pool, pool_global, lo, hi, bare_last, sc, p);
            __syncthreads();
        }
        if (phi > hi) break;      // this pass also holds later instructions: revisit it
    }
    *pcur = p;
}

// ---------------------------------------------------------------------------------------------------------------
// rare (jump event) paths of the tile kernel: kept out of line so that they do not set its register budget
// ---------------------------------------------------------------------------------------------------------------
static __device__ __noinline__ void event_apply(cplx* sm, const DevProgram& P, const Segment* sg, int instr, int k, double scale) {
    const int K = P.K;
    const Instr in = P.instrs[instr];
    const Channel ch = P.chans[in.channel];
    const cplx* M = reinterpret_cast<const cplx*>(P.pool + ch.kraus) + k * ch.dim * ch.dim;
    const int p0 = tile_pos(*sg, K, in.q0), p1 = in.q1 >= 0 ? tile_pos(*sg, K, in.q1) : -1;
    tile_apply_dense(sm, K, threadIdx.x, blockDim.x, ch.dim, p0, p1, M, scale);
    __syncthreads();
}

// reduced density matrix of tile positions (p0, p1) -> dst[32] (lower triangle), block-wide
static __device__ __noinline__ void event_reduce(const cplx* sm, int K, int p0, int p1, double* s_buf, double* s_out, double* dst) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    if (p1 < 0) {
        double acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0;
        reduce_rho_thread<2>(sm, K, tid, nthreads, p0, -1, acc);
        double v[6] = {acc[0], acc[1], acc[4], acc[5], acc[6], acc[7]};   // (0,0) (1,0) (1,1)
        block_reduce<6>(v, s_buf, s_out);
        if (tid == 0) {
            dst[0] = s_out[0]; dst[1] = s_out[1]; dst[4] = s_out[2]; dst[5] = s_out[3]; dst[6] = s_out[4]; dst[7] = s_out[5];
        }
    } else {
        double acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.0;
        reduce_rho_thread<4>(sm, K, tid, nthreads, p0, p1, acc);
        double v[20];
        int n = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) { v[n++] = acc[2 * (r * 4 + c)]; v[n++] = acc[2 * (r * 4 + c) + 1]; }
        block_reduce<20>(v, s_buf, s_out);
        if (tid == 0) {
            int m = 0;
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c <= r; ++c) { dst[2 * (r * 4 + c)] = s_out[m++]; dst[2 * (r * 4 + c) + 1] = s_out[m++]; }
        }
    }
    __syncthreads();
}

// Branch decision by the CTA that delivered the LAST partial of a state (generated event kernels: no decide launch).
// src: the state's ntiles x RED_VALUES partials (written by other CTAs: read around L1); the sum runs in a fixed order --
// warp w adds the tiles o = w (mod nwarps), warp 0 adds the warps' sums in order -- so it does not depend on which CTA
// happens to be last.  s_buf: >= (nwarps + 1) * 32 doubles of shared memory.
static __device__ __noinline__ void fused_decide(const DevProgram& P, const double* src, int ntiles, int instr, double u,
                                                 Decision* out, double* s_buf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // eight independent accumulators per lane (the loads are L2 round trips: a single dependent chain took ~60 us for 512
    // tiles and sat in the launch's tail), added in a fixed order
    double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int o = warp; o < ntiles; o += 8 * nwarps) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int oo = o + j * nwarps;
            if (oo < ntiles) acc[j] += __ldcg(src + (long long)oo * RED_VALUES + lane);
        }
    }
    const double s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    s_buf[warp * RED_VALUES + lane] = s;
    __syncthreads();
    if (warp == 0) {
        double* rho = s_buf + nwarps * RED_VALUES;
        double t = 0.0;
        for (int w = 0; w < nwarps; ++w) t += s_buf[w * RED_VALUES + lane];
        rho[lane] = t;
        __syncwarp();
        if (lane == 0) {
            const Channel ch = P.chans[P.instrs[instr].channel];
            mirror_rho(rho, ch.dim);
            double scale;
            out->k = decide_branch(rho, ch, P.pool, u, &scale);
            out->scale = scale;
        }
    }
    __syncthreads();
}

// decide_branch with the block's threads: thread k computes p_k = Re tr(M_k rho) (the same expression, term by term, as
// decide_branch), thread 0 adds them up and picks in order.  rho: shared memory (mirrored), s_p: >= ch.count doubles of shared
// memory.  Result in *k_out / *scale_out (shared) after the closing barrier.  Serial fall-back when s_p is too small.
static __device__ __noinline__ void decide_branch_block(const double* rho, const Channel& ch, const double* __restrict__ pool, double u,
                                                        double* s_p, int s_p_len, int* k_out, double* scale_out) {
    const int tid = threadIdx.x;
    if (ch.count > s_p_len) {
        if (tid == 0) { double sc; *k_out = decide_branch(rho, ch, pool, u, &sc); *scale_out = sc; }
        __syncthreads();
        return;
    }
    const int d = ch.dim, dd = d * d;
    const double* mk = pool + ch.mk;
    for (int k = tid; k < ch.count; k += blockDim.x) {
        double p = 0.0;
        for (int r = 0; r < d; ++r)
            for (int c = 0; c < d; ++c)
                p += mk[2 * (k * dd + r * d + c)] * rho[2 * (c * d + r)] - mk[2 * (k * dd + r * d + c) + 1] * rho[2 * (c * d + r) + 1];
        s_p[k] = p;
    }
    __syncthreads();
    if (tid == 0) {
        double sum = 0.0;
        for (int k = 0; k < ch.count; ++k) sum += s_p[k] > 0.0 ? s_p[k] : 0.0;
        int pick = ch.count - 1;
        double cp = 0.0, ppick = 0.0;
        for (int k = 0; k < ch.count; ++k) {
            double p = s_p[k];
            if (p < 0.0) p = 0.0;
            cp += p / sum;
            ppick = p;
            if (k == ch.count - 1 || !(cp < u)) { pick = k; break; }
        }
        const double st = pool[ch.scale + pick];
        *scale_out = ppick > 0.0 ? 1.0 / (st * sqrt(ppick)) : 0.0;
        *k_out = pick;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// tile kernel
// ---------------------------------------------------------------------------------------------------------------
// Persistent: gridDim.x CTAs walk the launch's (work, tile) pairs with stride gridDim.x.  Shared memory holds the
// segment's program blob and TWO tile buffers: while tile i is processed, tile i+1 streams in through cp.async.
template <int R, int FEAT, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
tile_kernel(DevProgram P, const __grid_constant__ SegConst sc, const unsigned char* __restrict__ blob_g, int blob_bytes, int blob_cap, int nbuf,
            const Work* __restrict__ works, int nworks,
            cplx* __restrict__ states, int tile_shift, double* __restrict__ red_part, double* __restrict__ norm_part,
            const Decision* __restrict__ decisions) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_buf[32 * 20];
    __shared__ double s_out[20];
    const int K = P.K, tid = threadIdx.x, nthreads = blockDim.x;
    const int ntiles = 1 << tile_shift;
    const long long total = (long long)nworks << tile_shift;
    // shared memory: [blob (blob_cap bytes)] [nbuf tile buffers] [jb table]
    cplx* const buf0 = reinterpret_cast<cplx*>(smem_raw + blob_cap);
    const int bufmask = nbuf - 1;          // nbuf is 1 or 2

    // ---- stage the segment program ----
    for (int i = tid * 16; i < blob_bytes; i += nthreads * 16) cp_async16(smem_raw + i, blob_g + i);
    cp_async_wait_all();
    __syncthreads();
    const SegBlob* blob = reinterpret_cast<const SegBlob*>(smem_raw);
    const Segment* sg = &blob->seg;
    const Pass* passes = reinterpret_cast<const Pass*>(smem_raw + blob->pass_off);
    const Op* ops = reinterpret_cast<const Op*>(smem_raw + blob->op_off);
    const double* pool = reinterpret_cast<const double*>(smem_raw + blob->pool_off);
    const int npass = blob->npass;

    const unsigned offA = apply_runs(sg->eruns, (unsigned)tid);
    const int sA = swz(tid);
    const unsigned limit = 1u << P.nbits;
    // per-thread base tile index of every pass: tile independent, computed once per CTA
    unsigned short* jb_tab = reinterpret_cast<unsigned short*>(buf0 + ((size_t)nbuf << K));
    const int jb_cap = npass < JB_TABLE_PASSES ? npass : JB_TABLE_PASSES;
    for (int p = 0; p < jb_cap; ++p) jb_tab[p * nthreads + tid] = (unsigned short)apply_runs(passes[p].truns, (unsigned)tid);

    auto issue_load = [&](long long t, cplx* sm, const Work& w) {
        if (w.flags & WF_INIT) return;                           // |0..0> is written by the processing step
        const unsigned base = apply_runs(sg->bruns, (unsigned)(t & (ntiles - 1)));
        const cplx* st = states + ((long long)w.src << P.nbits) + base;
#pragma unroll
        for (int it = 0; it < (1 << R); ++it) {
            const unsigned off = offA | sg->it_off[it];
            cplx* dst = sm + (sA ^ sg->it_swz[it]);
            if (off < limit) cp_async16(dst, st + off);
            else *dst = cplx{0.0, 0.0};
        }
    };

    long long t = blockIdx.x;
    if (t >= total) return;
    Work w = works[t >> tile_shift];
    issue_load(t, buf0, w);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (int iter = 0; t < total; ++iter, t += gridDim.x) {
        cplx* sm = buf0 + ((size_t)(iter & bufmask) << K);
        const long long tn = t + gridDim.x;
        Work wn = w;
        if (nbuf == 2) {
            if (tn < total) {
                wn = works[tn >> tile_shift];
                issue_load(tn, buf0 + ((size_t)((iter + 1) & 1) << K), wn);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");      // everything but the newest group has landed
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        __syncthreads();
        const int o = (int)(t & (ntiles - 1));
        if (w.flags & WF_INIT) {
#pragma unroll
            for (int it = 0; it < (1 << R); ++it) sm[sA ^ sg->it_swz[it]] = cplx{0.0, 0.0};
            if (o == 0 && tid == 0) sm[swz(0)] = cplx{1.0, 0.0};
            __syncthreads();
        }
        // ---- jump operator of the event that ended the previous piece ----
        if (w.flags & WF_PRO_STATIC) event_apply(sm, P, sg, w.ev_instr, w.ev_k, 1.0);
        else if (w.flags & WF_PRO_DYNAMIC) event_apply(sm, P, sg, w.ev_instr, decisions[w.slot].k, decisions[w.slot].scale);
        // ---- the segment's passes; a jump known from its uniform alone is applied on the way (WF_MID) ----
        {
            int pcur = 0, lo = w.instr_lo;
            for (int part = (w.flags & WF_MID) ? 0 : 1; part < 2; ++part) {
                const int hi = part == 0 ? w.mid_instr + 1 : w.instr_hi;
                const bool bare = part == 0 ? true : (w.flags & WF_BARE_LAST) != 0;
                run_range<R, FEAT, true>(sm, K, passes, ops, pool, P.pool, &pcur, npass, lo, hi, bare, sc, jb_tab, jb_cap);
                if (part == 0) {
                    event_apply(sm, P, sg, w.mid_instr, (int)((unsigned)w.flags >> 16), 1.0);
                    lo = w.mid_instr + 1;
                }
            }
        }
        // ---- reductions ----
        if (w.flags & WF_REDUCE) {
            const Instr in = P.instrs[w.instr_hi - 1];
            const int p0 = tile_pos(*sg, K, in.q0), p1 = in.q1 >= 0 ? tile_pos(*sg, K, in.q1) : -1;
            event_reduce(sm, K, p0, p1, s_buf, s_out, red_part + ((long long)w.slot * ntiles + o) * RED_VALUES);
        }
        if (w.flags & WF_NORM) {
            double v[1] = {0.0};
            for (int i = tid; i < (1 << K); i += nthreads) v[0] += cnorm2(sm[i]);
            block_reduce<1>(v, s_buf, s_out);
            if (tid == 0) norm_part[(long long)w.slot * ntiles + o] = s_out[0];
        }
        // ---- store ----
        {
            const unsigned base = apply_runs(sg->bruns, (unsigned)o);
            cplx* st = states + ((long long)w.slot << P.nbits) + base;
#pragma unroll
            for (int it = 0; it < (1 << R); ++it) {
                const unsigned off = offA | sg->it_off[it];
                if (off < limit) st_stream(st + off, sm[sA ^ sg->it_swz[it]]);
            }
        }
        __syncthreads();          // the buffer is refilled by the load issued in the next iteration
        if (nbuf == 1 && tn < total) {
            wn = works[tn >> tile_shift];
            issue_load(tn, buf0, wn);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        w = wn;
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// direct kernel: the sweep of the works that run a WHOLE segment without any jump event (most of them)
// ---------------------------------------------------------------------------------------------------------------
// Same persistent walk and tile geometry as tile_kernel, lean operator set, no event / reduction code.  Two things
// differ: the LAST pass stores its amplitudes from registers straight to HBM (64-byte runs per 4 lanes: the low state
// bits are the low thread-index bits of that pass, SegConst::direct) instead of going back through shared memory, and
// as soon as the last pass holds its amplitudes in registers the tile buffer is refilled with the CTA's next tile
// (cp.async), so that load overlaps the last pass' arithmetic and stores.  Works: slot (written), src (read).
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
direct_kernel(DevProgram P, const __grid_constant__ SegConst sc, const unsigned char* __restrict__ blob_g, int blob_bytes, int blob_cap,
              const Work* __restrict__ works, int nworks, cplx* __restrict__ states, int tile_shift) {
    constexpr int R = 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = P.K, tid = threadIdx.x, nthreads = blockDim.x;
    const int ntiles = 1 << tile_shift;
    const long long total = (long long)nworks << tile_shift;
    cplx* const sm = reinterpret_cast<cplx*>(smem_raw + blob_cap);
    for (int i = tid * 16; i < blob_bytes; i += nthreads * 16) cp_async16(smem_raw + i, blob_g + i);
    cp_async_wait_all();
    __syncthreads();
    const SegBlob* blob = reinterpret_cast<const SegBlob*>(smem_raw);
    const Segment* sg = &blob->seg;
    const Pass* passes = reinterpret_cast<const Pass*>(smem_raw + blob->pass_off);
    const Op* ops = reinterpret_cast<const Op*>(smem_raw + blob->op_off);
    const double* pool = reinterpret_cast<const double*>(smem_raw + blob->pool_off);
    const int npass = sc.npass;
    // per-thread constants of the CTA's lifetime: base tile index of every pass, state offset of the last pass
    unsigned short* jb_tab = reinterpret_cast<unsigned short*>(sm + ((size_t)1 << K));
    for (int p = 0; p < npass; ++p) jb_tab[p * nthreads + tid] = (unsigned short)apply_runs(passes[p].truns, (unsigned)tid);
    const unsigned gj_last = apply_runs(sc.gruns_last, (unsigned)tid);
    const unsigned offA = apply_runs(sg->eruns, (unsigned)tid);
    const int sA = swz(tid);
    const int lo = sg->instr_lo, hi = sg->instr_hi;
    auto issue_load = [&](long long tt) {
        const cplx* st = states + ((long long)works[tt >> tile_shift].src << P.nbits) + apply_runs(sg->bruns, (unsigned)(tt & (ntiles - 1)));
#pragma unroll
        for (int it = 0; it < (1 << R); ++it) cp_async16(sm + (sA ^ sg->it_swz[it]), st + (offA | sg->it_off[it]));
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    long long t = blockIdx.x;
    if (t >= total) return;
    issue_load(t);
    for (; t < total; t += gridDim.x) {
        const long long tn = t + gridDim.x;
        cplx* const dst = states + ((long long)works[t >> tile_shift].slot << P.nbits) + apply_runs(sg->bruns, (unsigned)(t & (ntiles - 1)));
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
        for (int p = 0; p < npass; ++p) {
            const bool last = p == npass - 1;
            auto refill = [&]() {
                if (!last) return;
                __syncthreads();             // every thread holds its amplitudes: the buffer is free
                if (tn < total) issue_load(tn);
            };
            run_pass_thread<R, 0, true, decltype(refill), true>(sm, K, (int)jb_tab[p * nthreads + tid], passes[p], ops, pool, P.pool, lo, hi, false,
                                                                sc, p, last ? dst : nullptr, gj_last, refill);
            if (!last) __syncthreads();
        }
    }
}

#ifndef NCB_TEMPLATES_ONLY   // the small non-template kernels live in ncb_engine.cu's module only
// ---------------------------------------------------------------------------------------------------------------
// decide kernel: one warp per state-dependent event
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
decide_kernel(DevProgram P, const DecideItem* __restrict__ items, int nitems, const double* __restrict__ red_part,
              int ntiles, Decision* __restrict__ decisions) {
    // one 256-thread block per event: lane e owns value e of the 32-value partial, warp w sums the tiles o = w (mod 8);
    // the 8 warp partials are then added in warp order, so the result does not depend on scheduling
    __shared__ double part[8][RED_VALUES];
    __shared__ double rho[RED_VALUES];
    const int it = blockIdx.x;
    if (it >= nitems) return;
    const DecideItem d = items[it];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* src = red_part + (long long)d.slot * ntiles * RED_VALUES;
    // eight independent accumulators per lane, added in a fixed order: the loads are L2 round trips and a single dependent
    // chain over 64 tiles per warp made this tiny kernel 75 us long
    double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int o = warp; o < ntiles; o += 64) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int oo = o + 8 * j;
            if (oo < ntiles) acc[j] += src[(long long)oo * RED_VALUES + lane];
        }
    }
    const double s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    part[warp][lane] = s;
    __syncthreads();
    const Channel ch = P.chans[P.instrs[d.instr].channel];
    if (warp == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += part[w][lane];
        rho[lane] = t;
        __syncwarp();
        if (lane == 0) mirror_rho(rho, ch.dim);
    }
    __syncthreads();
    // p_k = tr(M_k rho) by thread k (Eagle's two-qubit channels keep up to 144 operators: a single lane walking them made
    // this launch -- it sits between the pieces of every state-dependent event -- 170 us long), the pick in order by thread 0
    __shared__ int s_k;
    __shared__ double s_scale;
    decide_branch_block(rho, ch, P.pool, d.u, &part[0][0], 8 * RED_VALUES, &s_k, &s_scale);
    if (threadIdx.x == 0) {
        decisions[d.slot].k = s_k;
        decisions[d.slot].scale = s_scale;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// generic k-bit unitary (5 <= k <= 8): gather -> shared-memory mat-vec -> scatter (QuantumGates.hpp:968-1064 takes any k;
// the register-blocked passes stop at 4 bits).  One thread per (group, row); a block works on 256 >> k groups at a time.
// ---------------------------------------------------------------------------------------------------------------
struct GenericArgs {
    int nb;
    int bits[8];               // matrix bit i <-> state bit bits[i]
};
__global__ void __launch_bounds__(256)
generic_unitary_kernel(const Work* __restrict__ works, int nworks, cplx* __restrict__ states, int nbits, GenericArgs ga,
                       const cplx* __restrict__ U) {
    __shared__ cplx g_in[256];
    const int k = ga.nb, dimk = 1 << k, gpb = 256 >> k;
    const int lg = threadIdx.x >> k, row = threadIdx.x & (dimk - 1);
    const long long ngroups = 1ll << (nbits - k), total = (long long)nworks * ngroups;
    unsigned tmask = 0;
    for (int i = 0; i < k; ++i) tmask |= 1u << ga.bits[i];
    unsigned spread = 0;                                       // this thread's row, spread over the target bits
    for (int i = 0; i < k; ++i) spread |= (unsigned)((row >> i) & 1) << ga.bits[i];
    for (long long it = (long long)blockIdx.x * gpb; it < total; it += (long long)gridDim.x * gpb) {
        const long long gi = it + lg;
        const bool active = gi < total;
        Work w{};
        long long off = 0;
        if (active) {
            w = works[gi / ngroups];
            long long g = gi % ngroups;
            unsigned base = 0;                                  // deposit g into the non-target bits
            for (int b = 0; b < nbits; ++b)
                if (!((tmask >> b) & 1u)) { base |= (unsigned)(g & 1) << b; g >>= 1; }
            off = (long long)(base | spread);
            cplx v;
            if (w.flags & WF_INIT) v = cplx{off == 0 ? 1.0 : 0.0, 0.0};
            else v = states[((long long)w.src << nbits) + off];
            g_in[lg * dimk + row] = v;
        }
        __syncthreads();
        if (active) {
            const cplx* Ur = U + (size_t)row * dimk;
            const cplx* x = g_in + lg * dimk;
            cplx acc = cmul(Ur[0], x[0]);
            for (int c = 1; c < dimk; ++c) acc = cfma(Ur[c], x[c], acc);
            states[((long long)w.slot << nbits) + off] = acc;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// tail kernels
// ---------------------------------------------------------------------------------------------------------------
// scale[slot] = weight[slot] / sum_tiles norm_part[slot][tile]   (weight: how many trajectories end in this slot's state)
__global__ void norm_finalize_kernel(const double* __restrict__ norm_part, int ntiles, int nslots, const double* __restrict__ weight,
                                     double* __restrict__ scale) {
    const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (slot >= nslots) return;
    const double w = weight[slot];
    if (w == 0.0) { if (lane == 0) scale[slot] = 0.0; return; }      // unused slot: its partials are stale
    double s = 0.0;
    for (int o = lane; o < ntiles; o += 32) s += norm_part[(long long)slot * ntiles + o];
    s = warp_sum(s);
    if (lane == 0) scale[slot] = w / s;
}

// probs[j] += sum_slots |psi_slot[j]|^2 * scale[slot]   (slots in order: deterministic; slots with scale 0 are not read)
__global__ void accumulate_kernel(double* __restrict__ probs, const cplx* __restrict__ states, int nbits, int nslots,
                                  const double* __restrict__ scale) {
    const long long dim = 1ll << nbits;
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < dim; j += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int b = 0; b < nslots; ++b) {
            const double f = scale[b];
            if (f == 0.0) continue;
            const cplx a = states[((long long)b << nbits) + j];
            s += (a.x * a.x + a.y * a.y) * f;
        }
        probs[j] += s;
    }
}

// probs[j] += sum_c partial[c][j]
// trunk (optional): + n_trunk[0] * trunk[j], the trajectories without any jump (resident_kernel, ResidentArgs::share)
__global__ void sum_partials_kernel(double* __restrict__ probs, const double* __restrict__ partial, long long dim, int nparts,
                                    const double* __restrict__ trunk = nullptr, const unsigned long long* __restrict__ n_trunk = nullptr) {
    const double wt = trunk ? (double)n_trunk[0] : 0.0;
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < dim; j += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < nparts; ++c) s += partial[c * dim + j];
        if (wt != 0.0) s += wt * trunk[j];
        probs[j] += s;
    }
}

// probs[r] = Re rho[r][r] with vec index r | (r << nq)
__global__ void diag_kernel(double* __restrict__ probs, const cplx* __restrict__ rho, int nq) {
    const long long dim = 1ll << nq;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < dim; r += (long long)gridDim.x * blockDim.x)
        probs[r] = rho[r | (r << nq)].x;
}

__global__ void abs2_kernel(double* __restrict__ probs, const cplx* __restrict__ state, long long dim) {
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < dim; j += (long long)gridDim.x * blockDim.x)
        probs[j] = cnorm2(state[j]);
}

__global__ void scale_kernel(double* __restrict__ v, long long dim, double f) {
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < dim; j += (long long)gridDim.x * blockDim.x) v[j] *= f;
}

// one read-out error pass: 2x2 real matrix on probability pairs at stride 1 << q (MeasurementErrorApplicator.cpp:65-81)
__global__ void measurement_error_kernel(double* __restrict__ probs, long long half, int q, double m00, double m01, double m10, double m11) {
    const long long stride = 1ll << q;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < half; p += (long long)gridDim.x * blockDim.x) {
        const long long i = (p & (stride - 1)) | ((p & ~(stride - 1)) << 1), j = i | stride;
        const double p0 = probs[i], p1 = probs[j];
        probs[i] = m00 * p0 + m01 * p1;
        probs[j] = m10 * p0 + m11 * p1;
    }
}
// marginal over the qubits in keep_mask (m of them): out[i] = scale * sum over the other qubits, i little-endian over the
// kept qubits in ascending order (HelperFunctions.py:9-32; numpy sums the dropped axes, here one thread per output in
// ascending order of the dropped bits)
__global__ void marginal_kernel(double* __restrict__ out, const double* __restrict__ in, int nbits, unsigned keep_mask, int m, double scale) {
    const long long nout = 1ll << m, nrest = 1ll << (nbits - m);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nout; i += (long long)gridDim.x * blockDim.x) {
        long long base = 0;
        for (int b = 0, k = 0; b < nbits; ++b)
            if ((keep_mask >> b) & 1u) { base |= ((i >> k) & 1ll) << b; ++k; }
        double s = 0.0;
        for (long long r = 0; r < nrest; ++r) {
            long long off = 0;
            for (int b = 0, k = 0; b < nbits; ++b)
                if (!((keep_mask >> b) & 1u)) { off |= ((r >> k) & 1ll) << b; ++k; }
            s += in[base | off];
        }
        out[i] = s * scale;
    }
}

// little-endian -> big-endian index order over m bits (QuantumCircuit.py:445)
__global__ void bit_reverse_kernel(double* __restrict__ out, const double* __restrict__ in, int m) {
    const long long n = 1ll << m;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long r = 0;
        for (int b = 0; b < m; ++b) r |= ((i >> b) & 1ll) << (m - 1 - b);
        out[r] = in[i];
    }
}
#endif  // NCB_TEMPLATES_ONLY

// ---------------------------------------------------------------------------------------------------------------
// resident kernel: whole trajectory in one CTA's shared memory
// ---------------------------------------------------------------------------------------------------------------
struct ResidentArgs {
    long long traj_begin, traj_count;
    unsigned long long seed;
    const double* uniforms;     // NULL: Philox; else [traj_count][G]
    double* partial_probs;      // [gridDim.x][2^nbits] accumulated |psi|^2 (may be NULL)
    cplx* states_out;           // optional [traj_count][2^nbits]: final normalised states
    int normalize;              // divide by the norm (program is not norm preserving)
    unsigned long long* counters;   // [0] events, [1] state-dependent events, [2] trajectories without any event (share)
    // share: a trajectory whose draws all pick the base branch is the same evolution for everybody: block 0 computes it
    // once into trunk_probs (normalised |psi|^2), the others are only counted
    int share;
    double* trunk_probs;
};

// ---------------------------------------------------------------------------------------------------------------
// generated resident kernels (ncb_spec.cpp: ncb_res_trunk / ncb_res): arguments and the two planning kernels
// ---------------------------------------------------------------------------------------------------------------
// A trajectory is the jump-free evolution (the "trunk") up to its first flagged draw.  The trunk runs ONCE (ncb_res_trunk,
// one CTA) and leaves its state at every pass boundary in global memory (2^K amplitudes each: L2 resident); the planning
// kernel finds every trajectory's first flagged draw; the trajectories that leave the trunk are sorted by it (stable radix sort:
// deterministic, longest remaining work first, so that dealing them out round-robin balances the CTAs) and ncb_res runs each
// of them from the last checkpoint before that draw -- the others are only counted and weighted with the trunk's result.
struct ResCkpt { int32_t pass, instr; };    // checkpoint to start from (global pass index; -1: |0...0>) and the first instruction still to run
struct ResArgs {
    long long traj_begin;
    unsigned long long seed;
    const double* uniforms;          // NULL: Philox; else [traj_count][G], by execution position
    const double* thresh;            // [n_noisy][2]: a draw u with thresh[0] < u <= thresh[1] selects the base branch for sure
    const unsigned char* blobs;      // the program's segment blobs (phase tables and 2x2 coefficients are staged from their pools)
    const int32_t* seg_pool;         // [nseg][2]: byte offset of segment s's pool in blobs, bytes
    const uint32_t* item_first;      // [count] execution position of the first flagged draw, ascending (longest work first);
    const int32_t* item_traj;        // [count] its trajectory (index within the call); only the first n_items[0] entries are run
    const int32_t* n_items;
    int32_t count;                   // trajectories of this call
    const ResCkpt* ckpt_for;         // [G + 1], by execution position of the first flagged draw
    cplx* ckpt;                      // [npass + 1][2^K]: trunk states in the shared-memory (swizzled) layout; [npass] = final state
    double* partial;                 // [gridDim.x][2^nbits]
    cplx* states_out;                // optional [traj_count][2^nbits]
    double* trunk_probs;             // [2^nbits] normalised |psi|^2 of the trunk
    unsigned long long* counters;    // [0] events, [1] state-dependent events, [2] trajectories that never leave the trunk
    int32_t normalize, acc_global;
};

#ifndef NCB_TEMPLATES_ONLY
constexpr int RES_PLAN_THREADS = 128;
// One warp per trajectory: key[t] = execution position of trajectory t's first flagged draw (G: none; G + 1: none and not wanted
// as an item), val[t] = t; n_items counts the wanted ones.  The lanes test 32 draws at a time (noisy[] ascends, so the first
// set lane of the first non-empty ballot is the first flagged draw).
__global__ void __launch_bounds__(RES_PLAN_THREADS)
res_plan_kernel(const double* __restrict__ thresh, const int32_t* __restrict__ noisy, const int32_t* __restrict__ orig, int n_noisy, int G,
                unsigned long long seed, long long traj_begin, int count, const double* __restrict__ uniforms, int all,
                uint32_t* __restrict__ key, int32_t* __restrict__ val, int32_t* __restrict__ n_items) {
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * (RES_PLAN_THREADS / 32) + (threadIdx.x >> 5);
    if (t < count) {
        const double* utab = uniforms ? uniforms + (long long)t * G : nullptr;
        int first = G;
        for (int a0 = 0; a0 < n_noisy; a0 += 32) {
            const int a = a0 + lane;
            bool flagged = false;
            if (a < n_noisy) {
                const int i = noisy[a];
                const double u = utab ? utab[i] : philox_uniform(seed, (unsigned long long)(traj_begin + t), (uint32_t)orig[i]);
                flagged = !(u > thresh[2 * a] && u <= thresh[2 * a + 1]);
            }
            const unsigned m = __ballot_sync(0xffffffffu, flagged);
            if (m) { first = noisy[a0 + __ffs(m) - 1]; break; }
        }
        if (lane == 0) {
            const bool want = all || first < G;
            key[t] = (uint32_t)(want ? first : G + 1);
            val[t] = t;
            if (want) atomicAdd(&s_cnt, 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(n_items, s_cnt);
}

// probs[j] += sum_c partial[c][j] (+ n_trunk * trunk[j]) with 8 threads per column: thread (j, g) adds the partials g, g + 8, ...
// in order, thread (j, 0) the eight sums in order -- deterministic for a given number of partials
__global__ void __launch_bounds__(256)
sum_partials_wide_kernel(double* __restrict__ probs, const double* __restrict__ partial, long long dim, int nparts,
                         const double* __restrict__ trunk, const unsigned long long* __restrict__ n_trunk) {
    __shared__ double s[8][33];
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    const long long j = (long long)blockIdx.x * 32 + c;
    double v = 0.0;
    if (j < dim)
        for (int p = g; p < nparts; p += 8) v += partial[(long long)p * dim + j];
    s[g][c] = v;
    __syncthreads();
    if (g == 0 && j < dim) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += s[k][c];
        if (trunk) { const double wt = (double)n_trunk[0]; if (wt != 0.0) t += wt * trunk[j]; }
        probs[j] += t;
    }
}
#endif  // NCB_TEMPLATES_ONLY

template <int R, int FEAT, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
resident_kernel(DevProgram P, ResidentArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(smem_raw);                       // 2^K amplitudes
    double* acc = reinterpret_cast<double*>(sm + (1 << P.K));            // 2^nbits probabilities
    unsigned* evmask = reinterpret_cast<unsigned*>(acc + (1 << P.nbits));   // ceil(G/32) words
    __shared__ double s_buf[32 * 20];
    __shared__ double s_out[RED_VALUES];
    __shared__ double s_red[20];
    __shared__ int s_k;
    __shared__ double s_scale;
    const int K = P.K, tid = threadIdx.x, nthreads = blockDim.x, dim = 1 << P.nbits;
    const int nwords = (P.G + 31) >> 5;
    const Segment* sg = P.segs;     // every segment's tile is the whole state (identity map): the passes of all
                                    // segments form one list
    for (int j = tid; j < dim; j += nthreads) acc[j] = 0.0;
    unsigned long long n_events = 0, n_hard = 0;

    unsigned long long n_trunk = 0;
    // block 0 runs one extra iteration first (tl < 0): the jump-free trajectory
    for (long long tl = (long long)blockIdx.x - ((A.share && blockIdx.x == 0) ? (long long)gridDim.x : 0); tl < A.traj_count; tl += gridDim.x) {
        const bool is_trunk = tl < 0;
        const unsigned long long traj = (unsigned long long)(A.traj_begin + tl);
        const double* utab = A.uniforms && !is_trunk ? A.uniforms + tl * P.G : nullptr;
        // ---- plan: which instructions deviate from their base branch? ----
        for (int i = tid; i < nwords; i += nthreads) evmask[i] = 0u;
        for (int j = tid; j < (1 << K); j += nthreads) sm[j] = cplx{0.0, 0.0};
        __syncthreads();
        if (tid == 0) sm[swz(0)] = cplx{1.0, 0.0};
        int flagged = 0;
        for (int a = tid; a < P.n_noisy && !is_trunk; a += nthreads) {
            const int i = P.noisy[a];
            const Channel ch = P.chans[P.instrs[i].channel];
            const double* lo = P.pool + ch.bounds;
            const double* hi = lo + (ch.count - 1);
            const double u = utab ? utab[i] : philox_uniform(A.seed, traj, (uint32_t)P.orig[i]);
            const double left = ch.base == 0 ? -1.0 : hi[ch.base - 1] + EVENT_MARGIN;
            const double right = ch.base == ch.count - 1 ? 2.0 : lo[ch.base] - EVENT_MARGIN;
            if (!(u > left && u <= right)) { atomicOr(&evmask[i >> 5], 1u << (i & 31)); flagged = 1; }
        }
        if (A.share && !is_trunk) {
            if (!__syncthreads_or(flagged)) { ++n_trunk; continue; }      // block 0's trunk stands for this trajectory
        } else {
            __syncthreads();
        }
        // ---- run ----
        int cur = 0, pcur = sg->pass_begin;
        while (true) {
            // next flagged instruction >= cur
            int e = P.G;
            for (int wd = cur >> 5; wd < nwords; ++wd) {
                unsigned m = evmask[wd];
                if (wd == (cur >> 5)) m &= ~0u << (cur & 31);
                if (m) { e = (wd << 5) + __ffs(m) - 1; break; }
            }
            const bool ev = e < P.G;
            run_range<R, FEAT, false>(sm, K, P.passes, P.ops, P.pool, P.pool, &pcur, P.npass, cur, ev ? e + 1 : P.G, ev, *reinterpret_cast<const SegConst*>(P.pool));
            if (!ev) break;
            const Instr in = P.instrs[e];
            const Channel ch = P.chans[in.channel];
            const double u = utab ? utab[e] : philox_uniform(A.seed, traj, (uint32_t)P.orig[e]);
            int kmin, kmax;
            classify_uniform(P.pool + ch.bounds, P.pool + ch.bounds + (ch.count - 1), ch.count, u, &kmin, &kmax);
            int k = kmin;
            double scale = 1.0;
            bool base_after_all = false;
            if (kmin != kmax) {
                ++n_hard;
                event_reduce(sm, K, in.q0, ch.dim == 4 ? in.q1 : -1, s_buf, s_red, s_out);
                if (tid == 0) {
                    mirror_rho(s_out, ch.dim);
                    double sc;
                    s_k = decide_branch(s_out, ch, P.pool, u, &sc);
                    s_scale = sc;
                }
                __syncthreads();
                k = s_k; scale = s_scale;
            } else if (k == ch.base) {
                base_after_all = true;     // flagged by the cheap test only (u within the margin): base branch
            }
            if (!base_after_all) ++n_events;
            const cplx* M = reinterpret_cast<const cplx*>(P.pool + ch.kraus) + k * ch.dim * ch.dim;
            tile_apply_dense(sm, K, tid, nthreads, ch.dim, in.q0, in.q1, M, scale);
            __syncthreads();
            cur = e + 1;
        }
        // ---- accumulate ----
        double inv = 1.0;
        if (A.normalize) {
            double v[1] = {0.0};
            for (int j = tid; j < (1 << K); j += nthreads) v[0] += cnorm2(sm[j]);
            block_reduce<1>(v, s_buf, s_out);
            inv = 1.0 / s_out[0];
            __syncthreads();
        }
        if (is_trunk) {
            for (int j = tid; j < dim; j += nthreads) A.trunk_probs[j] = cnorm2(sm[swz(j)]) * inv;
        } else if (A.partial_probs) {
            for (int j = tid; j < dim; j += nthreads) acc[j] += cnorm2(sm[swz(j)]) * inv;
        }
        if (A.states_out) {
            const double f = sqrt(inv);
            for (int j = tid; j < dim; j += nthreads) {
                cplx a = sm[swz(j)];
                a.x *= f; a.y *= f;
                A.states_out[tl * dim + j] = a;
            }
        }
        __syncthreads();
    }
    if (A.partial_probs)
        for (int j = tid; j < dim; j += nthreads) A.partial_probs[(long long)blockIdx.x * dim + j] = acc[j];
    if (tid == 0 && A.counters) {
        atomicAdd(&A.counters[0], n_events);
        atomicAdd(&A.counters[1], n_hard);
        atomicAdd(&A.counters[2], n_trunk);
    }
}

}  // namespace ncb
