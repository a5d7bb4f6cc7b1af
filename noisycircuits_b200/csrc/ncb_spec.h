// Source generator of the specialised sweep kernels (see ncb_spec.cuh).
#pragma once
#include <string>

#include "ncb_compile.h"

namespace ncb {

// true when segment `seg` of P can run in a specialised kernel (lean program, tiled, valid constant-bank program with
// direct store)
bool spec_supported(const Program& P, int seg);
// CUDA source of `extern "C" __global__ void ncb_seg_<seg>(...)` for segments [seg_begin, seg_end) of P
std::string spec_source(const Program& P, int seg_begin, int seg_end, int nbuf = 2);
// bytes of segment seg's phase-table pool (the only part of the staged blob a specialised kernel needs) and its offset
void spec_pool(const Program& P, int seg, int* offset, int* bytes);
// resident CTAs per SM the generated kernels are compiled for (launch bounds)
int spec_ctas_per_sm(const Program& P, int nbuf);
// content hash of everything the generated code depends on (cache key)
std::string spec_key(const Program& P, int nbuf = 2);

}  // namespace ncb
