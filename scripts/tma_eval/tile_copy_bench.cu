// Evaluation of Blackwell's bulk asynchronous copies (cp.async.bulk + mbarrier, SASS UBLKCP) for the tile movement of
// the sweep kernels, against what they use today (cp.async 16-byte gathers into an XOR-swizzled tile, LDGSTS) and against
// plain register loads.  A "sweep with no operators": persistent CTAs walk the tiles of a batch of 2^n-amplitude states;
// a tile is 2^11 complex128 amplitudes (32 KB) made of runs of RUN contiguous bytes (64 B .. 4 KB, what the low state
// bits of the segment's tile give); every tile is read once and written once.
//
//   mode 0  cp.async.cg 16 B per thread and amplitude (8 per thread) -> shared memory -> ld.shared -> st.global.cs
//   mode 1  cp.async.bulk.shared.global per run, completion on an mbarrier (double buffered) -> ld.shared -> st.global.cs
//   mode 2  like 1, and the tile is written back with cp.async.bulk.global.shared (bulk_group) from shared memory
//   mode 3  ld.global.cs into registers -> st.global.cs (no shared memory)
//
// build: nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -o tile_copy_bench tile_copy_bench.cu
// usage: tile_copy_bench [trajectories=128] [reps=5]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

struct alignas(16) cplx { double x, y; };
constexpr int NBITS = 20, K = 11, THREADS = 256, TILE = 1 << K;   // amplitudes

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int swz(int j) { return j ^ (((j >> 3) ^ (j >> 6) ^ (j >> 9) ^ (j >> 12)) & 7); }

// state offset (in amplitudes) of run r of tile t: the run covers the low `lb` bits; the tile's other K - lb bits are
// the state bits [lb + gap, lb + gap + K - lb) -- `gap` non-tile bits in between, the rest of the tile number above
__device__ __forceinline__ long long run_offset(int t, int r, int lb, int gap) {
    const int hb = K - lb;                                       // tile bits above the run
    const long long lowgap = t & ((1 << gap) - 1);               // tile-number bits that fill the gap
    const long long high = t >> gap;                             // ... and the bits above the tile
    return (lowgap << lb) | ((long long)r << (lb + gap)) | (high << (lb + gap + hb));
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 3)
copy_kernel(const cplx* __restrict__ src, cplx* __restrict__ dst, long long ntiles_total, int lb, int gap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* buf = reinterpret_cast<cplx*>(smem_raw);               // two tiles
    __shared__ __align__(8) unsigned long long mbar[2];
    const int tid = threadIdx.x;
    const int run_amps = 1 << lb, nruns = TILE >> lb;
    const int tiles_per_state = 1 << (NBITS - K);
    if (MODE == 3) {
        for (long long t = blockIdx.x; t < ntiles_total; t += gridDim.x) {
            const long long base = (t / tiles_per_state) << NBITS;
            const int tt = (int)(t % tiles_per_state);
            cplx a[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int j = tid + m * THREADS;
                const long long off = base + run_offset(tt, j >> lb, lb, gap) + (j & (run_amps - 1));
                asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(a[m].x), "=d"(a[m].y) : "l"(src + off));
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int j = tid + m * THREADS;
                const long long off = base + run_offset(tt, j >> lb, lb, gap) + (j & (run_amps - 1));
                asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(dst + off), "d"(a[m].x), "d"(a[m].y) : "memory");
            }
        }
        return;
    }
    if (MODE >= 1) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    auto issue = [&](long long t, int b) {
        const long long base = (t / tiles_per_state) << NBITS;
        const int tt = (int)(t % tiles_per_state);
        cplx* sm = buf + (size_t)b * TILE;
        if (MODE == 0) {
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int j = tid + m * THREADS;
                const long long off = base + run_offset(tt, j >> lb, lb, gap) + (j & (run_amps - 1));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sm + swz(j))), "l"(src + off) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            if (tid == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[b])), "r"(TILE * 16) : "memory");
            __syncwarp();
            // one bulk copy per run, spread over the threads of the CTA (thread 0 has armed the barrier: the order between
            // expect_tx and complete_tx does not matter to the transaction count)
            for (int r = tid; r < nruns; r += THREADS) {
                const long long off = base + run_offset(tt, r, lb, gap);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(sm + (size_t)r * run_amps)), "l"(src + off), "r"(run_amps * 16), "r"(smem_u32(&mbar[b])) : "memory");
            }
        }
    };
    long long t = blockIdx.x;
    if (t >= ntiles_total) return;
    issue(t, 0);
    unsigned phase[2] = {0u, 0u};
    for (int it = 0; t < ntiles_total; t += gridDim.x, ++it) {
        const int b = it & 1;
        const long long tn = t + gridDim.x;
        if (MODE == 2 && tn < ntiles_total) {
            // the other buffer may still be read by the bulk store of the previous iteration
            if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
        }
        if (tn < ntiles_total) issue(tn, b ^ 1);
        else if (MODE == 0) asm volatile("cp.async.commit_group;" ::: "memory");
        if (MODE == 0) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncthreads();
        } else {
            asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra WAIT_%=;\n}" ::"r"(smem_u32(&mbar[b])), "r"(phase[b]) : "memory");
            phase[b] ^= 1u;
        }
        const long long base = (t / tiles_per_state) << NBITS;
        const int tt = (int)(t % tiles_per_state);
        cplx* sm = buf + (size_t)b * TILE;
        if (MODE == 2) {
            // (a real sweep would write its results into the tile here and make them visible to the async proxy)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            for (int r = tid; r < nruns; r += THREADS) {
                const long long off = base + run_offset(tt, r, lb, gap);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(smem_u32(sm + (size_t)r * run_amps)), "r"(run_amps * 16) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        } else {
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int j = tid + m * THREADS;
                const cplx v = sm[MODE == 0 ? swz(j) : j];
                const long long off = base + run_offset(tt, j >> lb, lb, gap) + (j & (run_amps - 1));
                asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(dst + off), "d"(v.x), "d"(v.y) : "memory");
            }
            __syncthreads();          // the buffer is refilled two iterations later, by the load issued in the next one
        }
    }
    if (MODE == 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (MODE == 0) asm volatile("cp.async.wait_all;" ::: "memory");
}

template <int MODE>
static float run(const cplx* src, cplx* dst, long long ntiles, int lb, int gap, int reps, int ctas) {
    const size_t smem = 2 * TILE * sizeof(cplx);
    cudaFuncSetAttribute(copy_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    copy_kernel<MODE><<<ctas, THREADS, MODE == 3 ? 0 : smem>>>(src, dst, ntiles, lb, gap);       // warm-up
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) copy_kernel<MODE><<<ctas, THREADS, MODE == 3 ? 0 : smem>>>(src, dst, ntiles, lb, gap);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (cudaGetLastError() != cudaSuccess) return -1.f;
    return ms / reps;
}

__global__ void fill(cplx* p, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = cplx{(double)(i & 1023), -(double)(i >> 10)};
}
__global__ void diff(const cplx* a, const cplx* b, long long n, unsigned long long* bad) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (a[i].x != b[i].x || a[i].y != b[i].y) atomicAdd(bad, 1ull);
}

int main(int argc, char** argv) {
    const int traj = argc > 1 ? atoi(argv[1]) : 128, reps = argc > 2 ? atoi(argv[2]) : 5;
    const long long n = (long long)traj << NBITS, ntiles = n >> K;
    cplx *src, *dst;
    unsigned long long* bad;
    if (cudaMalloc(&src, n * sizeof(cplx)) != cudaSuccess || cudaMalloc(&dst, n * sizeof(cplx)) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&bad, 8);
    fill<<<2048, 256>>>(src, n);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int ctas = prop.multiProcessorCount * 3;
    printf("{\"device\": \"%s\", \"trajectories\": %d, \"state_gb\": %.2f, \"tile_bytes\": %d, \"ctas\": %d, \"results\": [\n", prop.name, traj, n * 16.0 / 1e9, TILE * 16, ctas);
    const char* names[4] = {"cp.async 16 B gathers (LDGSTS), swizzled tile", "cp.async.bulk per run + mbarrier (UBLKCP), ld.shared + st.global",
                            "cp.async.bulk load and bulk store per run", "ld.global -> registers -> st.global"};
    bool first = true;
    for (int lb : {2, 3, 4, 6, 8}) {
        const int gap = 3;
        for (int mode = 0; mode < 4; ++mode) {
            cudaMemset(dst, 0, n * sizeof(cplx));
            float ms = -1;
            switch (mode) {
            case 0: ms = run<0>(src, dst, ntiles, lb, gap, reps, ctas); break;
            case 1: ms = run<1>(src, dst, ntiles, lb, gap, reps, ctas); break;
            case 2: ms = run<2>(src, dst, ntiles, lb, gap, reps, ctas); break;
            case 3: ms = run<3>(src, dst, ntiles, lb, gap, reps, ctas); break;
            }
            cudaMemset(bad, 0, 8);
            diff<<<2048, 256>>>(src, dst, n, bad);
            unsigned long long nbad = 0;
            cudaMemcpy(&nbad, bad, 8, cudaMemcpyDeviceToHost);
            printf("%s  {\"run_bytes\": %d, \"mode\": %d, \"how\": \"%s\", \"ms\": %.4f, \"gbs\": %.1f, \"mismatches\": %llu}", first ? "" : ",\n", 16 << lb, mode, names[mode], ms,
                   ms > 0 ? 2.0 * n * 16 / ms / 1e6 : 0.0, nbad);
            first = false;
            fflush(stdout);
        }
    }
    printf("\n]}\n");
    return 0;
}
