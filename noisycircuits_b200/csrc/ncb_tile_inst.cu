// One tile_kernel / resident_kernel variant per translation unit (compiled once per variant by the Makefile with
// -DNCB_INST_...): NVVM's inlining and ptxas' register allocation are decided per module, and the hot 64-register
// variant spills 5x more when it shares a module with the other variants.
#define NCB_TEMPLATES_ONLY
#include "ncb_kernels.cuh"

namespace ncb {
#if defined(NCB_INST_TILE)
template __global__ void tile_kernel<NCB_INST_R, NCB_INST_FEAT, NCB_INST_MAXT, NCB_INST_MINB>(
    DevProgram, const __grid_constant__ SegConst, const unsigned char*, int, int, int, const Work*, int, cplx*, int, double*, double*,
    const Decision*);
#elif defined(NCB_INST_DIRECT)
template __global__ void direct_kernel<NCB_INST_MAXT, NCB_INST_MINB>(DevProgram, const __grid_constant__ SegConst, const unsigned char*, int, int,
                                                                     const Work*, int, cplx*, int);
#else
template __global__ void resident_kernel<NCB_INST_R, NCB_INST_FEAT, NCB_INST_MAXT, NCB_INST_MINB>(DevProgram, ResidentArgs);
#endif
}  // namespace ncb
