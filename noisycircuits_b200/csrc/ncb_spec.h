// Source generator of the specialised sweep kernels (see ncb_spec.cuh).
#pragma once
#include <string>
#include <vector>

#include "ncb_compile.h"

namespace ncb {

struct SpecOptions {
    int nbuf = 2;     // shared-memory tile buffers of a whole-segment kernel (2: a tile's passes never wait for the previous
                      // tile's last reads)
    int load = 0;     // how a whole-segment kernel gets its tile: 0 staged in shared memory with cp.async (as the interpreting
                      // direct kernel does); 1 the first pass loads HBM -> registers (no staging, one barrier and one
                      // shared-memory round trip less); 2 like 1, and the next tile's loads are issued into a second register
                      // set when the last pass starts, so they overlap its arithmetic and stores
    int events = 1;   // 1: also generate ncb_seg_<k>_ev, the kernel of the works that carry a jump event (pieces of a segment)
    int factor = 1;   // 1 (literal mode): one real coefficient is factored out of every rotation and the product is multiplied
                      // back in once per segment -- half the FP64 instructions of the symmetric rotations, none for x
    int ctas = 0;     // > 0: resident CTAs per SM the kernels are compiled for (launch bounds; 0: from registers and shared memory)
    int ev_fast = 1;         // 1: in the event kernels a pass that lies entirely inside the range being run takes its fast list (the
                             // whole-segment kernel's text) instead of the guarded operators
    int resident_fast = 1;   // 1: the resident kernels run whole segments that hold no jump event through the passes' fast lists
    int literal = 1;  // 1: gate coefficients are written into the source as exact literals (ptxas materialises them next to
                      // their use; the cubin then belongs to this circuit AND these angles); 0: they are read from the
                      // SpecCoef kernel parameter (constant bank), so the source -- and the cubin cache key -- depends on the
                      // circuit's structure only and a parameter sweep compiles once (measured 13 % slower: ptxas hoists all
                      // constant loads out of the tile loop and pays for it in registers)
};

// true when segment `seg` of P can run in a specialised whole-segment kernel (lean program, tiled, valid constant-bank
// program with direct store)
bool spec_supported(const Program& P, int seg);
// true when the works of segment `seg` that carry jump events can run in a generated kernel (lean program, tiled)
bool spec_ev_supported(const Program& P, int seg);
// CUDA source of `extern "C" __global__ void ncb_seg_<seg>(...)` (and ncb_seg_<seg>_ev) for segments [seg_begin, seg_end)
// of P.  The source depends on the STRUCTURE of the program only (geometry, operator kinds, bit positions, table
// offsets): gate coefficients come in through a kernel parameter (constant bank), phase tables through the staged pool.
std::string spec_source(const Program& P, int seg_begin, int seg_end, const SpecOptions& opt = SpecOptions());
// bytes of segment seg's matrix pool (the only part of the staged blob a specialised kernel needs) and its offset
void spec_pool(const Program& P, int seg, int* offset, int* bytes);
// resident CTAs per SM the generated kernels are compiled for (launch bounds)
int spec_ctas_per_sm(const Program& P, const SpecOptions& opt);
// ---- resident programs (the whole state in one CTA's shared memory) ----
struct ResidentLayout {
    std::vector<int> pool_src, pool_len, pool_off;   // per segment: byte offset of its pool in Program::blobs, bytes (16-byte
                                                     // multiple), byte offset of its copy in the kernel's staged pool area
    int pool_bytes = 0;
    int acc_global = 0;   // 1: the CTA's probability accumulator lives in global memory (buys another CTA per SM)
    int ctas = 0;         // resident CTAs per SM (launch bounds)
    size_t smem = 0;      // dynamic shared memory of ncb_res (ncb_res_trunk uses the same amount)
};
bool spec_resident_supported(const Program& P);
ResidentLayout spec_resident_layout(const Program& P, const SpecOptions& opt = SpecOptions());
// [G + 1][2]: (checkpoint pass, first instruction still to run) by execution position of a trajectory's first flagged draw
std::vector<int32_t> spec_resident_ckpts(const Program& P);
// CUDA source of ncb_res_trunk / ncb_res ("" when the program is not resident or carries operators they do not cover)
// which: 0 both kernels, 1 ncb_res only, 2 ncb_res_trunk only (the engine compiles them as two translation units in parallel)
std::string spec_resident_source(const Program& P, const SpecOptions& opt = SpecOptions(), int which = 0);
// cheap hash of the records the generators read (no source is generated): equal hashes <=> the same kernels; used to see that
// an updated program (parameter sweep) keeps its loaded kernels.  A mismatch is re-checked with spec_key.
uint64_t spec_structure_hash(const Program& P, const SpecOptions& opt = SpecOptions());
// content hash of everything the generated code depends on (cache key): a hash of the source itself
std::string spec_key(const Program& P, const SpecOptions& opt = SpecOptions());

}  // namespace ncb
