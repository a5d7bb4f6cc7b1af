// Host-side compiler: (instruction list, Kraus tables) -> packed program of segments / passes / ops.
// Pure C++ (no CUDA calls): replaces the per-call parsing the reference does in
// Simulator.cpp:162-196 + NoiseParser.hpp:52-140 + FunctionMapper.hpp:25-79 by a one-time compilation.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "ncb_common.h"

namespace ncb {

enum Mode { MODE_MCWF = 0, MODE_DENSITY = 1, MODE_PURE = 2 };

// Gate opcodes of the C ABI (names of FunctionMapper.hpp:27-39)
enum Opcode { G_X = 0, G_SX, G_RZ, G_RX, G_RY, G_H, G_P, G_CZ, G_RZZ, G_ECR, G_CX, G_SWAP, G_UNITARY, G_COUNT };

struct CircuitView {               // mirrors ncb_circuit in include/ncb200.h
    int32_t num_qubits, num_instructions;
    const int32_t *opcode, *q0, *q1;
    const double* theta;
    const int32_t* channel;
    const int64_t* unitary_offset;
    const int32_t* unitary_nq;
    const int32_t* unitary_qubits;   // 8 per instruction, matrix bit b <-> unitary_qubits[8*g + b]
    const double* unitary_pool;      // interleaved re, im
    int32_t num_channels;
    const int64_t* channel_offset;   // in complex elements into kraus_pool
    const int32_t* channel_count;
    const int32_t* channel_dim;      // 2 or 4
    const double* kraus_pool;        // interleaved re, im, row-major operators
};

struct CompileOptions {
    int tile_bits = 11;     // K: amplitudes per shared-memory tile = 2^K (11: 32 KB tiles, 4 CTAs per SM)
    int resident_bits = 12; // states of up to 2^resident_bits amplitudes stay in one CTA's shared memory (resident kernel)
    int reg_bits = 3;       // R: amplitudes per thread = 2^R
    int low_bits = 2;       // the L lowest state bits are part of every tile (coalescing: runs of 2^L * 16 B = 64 B)
    int fuse = 1;           // 0: one instruction per segment (debug / ablation)
    int pass_order = 1;     // 1: inside a segment, instructions are ordered pass by pass (see compile_program)
    double blob_cost_scale = 1.0;   // multiplier of the segmenter's program-size estimate (raised automatically on overflow)
    int fuse_rotations = 1; // 1: commuting rx-like rotations of a pass are grouped and dispatched as one op (constant-bank fast list)
    int direct = 1;         // 1: first / last pass of a segment exchange their amplitudes with HBM directly where possible
    int decompose = 1;      // 1: controlled-type two-qubit gates (ecr, cx) run as one-qubit factors + a diagonal table
    int split_generic = 1;  // 1: in the fast lists a generic one-bit 2x2 hands its phases to neighbouring phase tables and runs as
                            // the rx-like operator that is left (compile_program: "D1 . X . D2")
    int reorder = 1;        // 1: execute commuting instructions in a tile-friendly (dependency-respecting) order;
                            // 0: strictly in program order
};

// A logical op in state-bit coordinates, before tiling
struct LogicalOp {
    int instr;
    int nb;                       // number of state bits
    int bits[8];                  // matrix bit i <-> state bit bits[i] (dense: <= 4 bits; merged diagonals: <= 6)
    bool diag;
    std::vector<double> mat;      // merged (2^nb x 2^nb complex, interleaved) or, if diag, 2^nb phases
    std::vector<double> bare;
    bool generic = false;         // dense operator on 5..8 bits (a user-supplied `unitary`): a segment of its own, run by
                                  // generic_unitary_kernel (gather -> shared-memory mat-vec -> scatter)
};

// A k-bit (5 <= k <= 8) dense unitary: does not fit the register-blocked passes (2^R amplitudes per thread, R <= 4).
struct GenericOp {
    int seg;                      // the segment that consists of this operator alone (Segment::pad[0] == 1)
    int nb;
    int bits[8];                  // matrix bit i <-> state bit bits[i]
    int mat;                      // offset (doubles) of the 2^nb x 2^nb row-major complex matrix in the program's pool
};

struct Program {
    int mode = MODE_MCWF;
    int nq = 0;            // qubits
    int nbits = 0;         // state bits (nq, or 2*nq for the density-matrix mode)
    int K = 0, R = 0;      // tile bits (may exceed nbits for tiny states: the extra bits are virtual), register bits
    int G = 0;
    bool has_dense4 = false;
    bool has_dense2 = false;
    double global_phase[2] = {1.0, 0.0};   // e^{i Phi}: product of the global phases factored out of one-qubit gates;
                                           // the engine's raw states are e^{-i Phi} x the true states
    bool norm_preserving = true;
    std::vector<double> pool;
    std::vector<Op> ops;
    std::vector<Pass> passes;
    std::vector<Segment> segs;
    std::vector<Instr> instrs;
    std::vector<Channel> chans;
    std::vector<GenericOp> generics;    // segments that are one 5..8-bit unitary (reference: apply_unitary_gate takes any k)
    std::vector<int32_t> draw_pos;      // position of the instruction's draw in the reference's MT stream, -1: none
    int num_draws = 0;
    std::vector<int32_t> noisy_instrs;  // execution positions that have a channel with >= 2 operators
    std::vector<int32_t> orig;          // orig[pos]: index in the caller's instruction list of the instruction executed
                                        // at position pos (identity unless reorder).  Inside the engine "instruction
                                        // index" always means the execution position; uniforms are keyed by orig[pos].
    std::vector<unsigned char> blobs;   // SegBlob of every segment, back to back (16-byte aligned)
    std::vector<int64_t> blob_off;      // byte offset of segment s's blob
    std::vector<SegConst> segconst;     // constant-bank part of every segment (kernel parameter of its launches)
    CompileOptions opt;
    int compile_attempts = 1;            // > 1: the segmenter's size estimate had to be raised (see compile_program)
};

// Throws std::runtime_error with a descriptive message on invalid input.
void compile_program(const CircuitView& c, int mode, const CompileOptions& opt, Program& out);

// Gate matrix in the reference's convention (2q index = bit(q0) + 2*bit(q1), QuantumGates.hpp:209-219); row-major,
// interleaved re/im.  dim = 2 or 4.
void gate_matrix(int opcode, double theta, double* out, int* dim);

// Events of one trajectory, in instruction order.
struct Event {
    int32_t instr;
    int32_t kmin, kmax;     // kmin == kmax: the branch is known (and differs from the channel's base branch)
    double u;
};
// uniform(instr) -> u in [0,1), instr = index in the caller's instruction list; Event::instr is an execution position
template <class F>
inline void plan_events(const Program& P, F&& uniform, std::vector<Event>& out) {
    out.clear();
    for (int32_t i : P.noisy_instrs) {
        const Channel& ch = P.chans[P.instrs[i].channel];
        const double* lo = P.pool.data() + ch.bounds;
        const double* hi = lo + (ch.count - 1);
        const double u = uniform(P.orig[i]);
        // fast path: inside the base branch's certain interval
        const double left = ch.base == 0 ? -1.0 : hi[ch.base - 1] + EVENT_MARGIN;
        const double right = ch.base == ch.count - 1 ? 2.0 : lo[ch.base] - EVENT_MARGIN;
        if (u > left && u <= right) continue;
        int kmin, kmax;
        classify_uniform(lo, hi, ch.count, u, &kmin, &kmax);
        if (kmin == kmax && kmin == ch.base) continue;
        out.push_back(Event{i, kmin, kmax, u});
    }
}

std::string describe_program(const Program& P);

}  // namespace ncb
