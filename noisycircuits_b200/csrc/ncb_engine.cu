// Runtime + C ABI of the B200 noisy-circuit engine (include/ncb200.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <atomic>
#include <mutex>
#include <dlfcn.h>
#include <dirent.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>
#include <fstream>
#include <sstream>
#include <vector>

#include "../../include/ncb200.h"
#include "ncb_compile.h"
#include "ncb_kernels.cuh"
#include "ncb_schedule.h"
#include "ncb_spec.h"
#include <cub/device/device_radix_sort.cuh>

namespace ncb {

// The kernel variants are instantiated one per translation unit (ncb_tile_inst.cu, see the Makefile).
#define NCB_EXTERN_TILE(R, F, T, B)                                                                                     \
    extern template __global__ void tile_kernel<R, F, T, B>(DevProgram, const __grid_constant__ SegConst, const unsigned char*, int, \
                                                            int, int, const Work*, int, cplx*, int, double*, double*,     \
                                                            const Decision*);
NCB_EXTERN_TILE(3, 0, 512, 1)
NCB_EXTERN_TILE(3, 0, 512, 2)
NCB_EXTERN_TILE(3, 0, 256, 3)
NCB_EXTERN_TILE(3, FEAT_DENSE2, 256, 1)
NCB_EXTERN_TILE(4, FEAT_DENSE2 | FEAT_DENSE4, 256, 1)
NCB_EXTERN_TILE(4, 0, 256, 2)
#undef NCB_EXTERN_TILE
extern template __global__ void direct_kernel<512, 2>(DevProgram, const __grid_constant__ SegConst, const unsigned char*, int, int, const Work*, int,
                                                     cplx*, int);
extern template __global__ void resident_kernel<3, 0, 512, 1>(DevProgram, ResidentArgs);
extern template __global__ void resident_kernel<3, FEAT_DENSE2, 256, 1>(DevProgram, ResidentArgs);
extern template __global__ void resident_kernel<4, FEAT_DENSE2 | FEAT_DENSE4, 256, 1>(DevProgram, ResidentArgs);

static thread_local std::string g_error;

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
#define CK(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess)                                                                                     \
            throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) CK(cudaFree(p));
        p = nullptr; cap = 0;
        CK(cudaMalloc(&p, n * sizeof(T)));
        cap = n;
    }
    void reserve_growing(size_t n) { if (n > cap) reserve(n + n / 4 + 64); }   // per-batch lists: sizes vary a few percent
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <class T>
struct PinnedBuf {
    T* p = nullptr;
    size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) CK(cudaFreeHost(p));
        p = nullptr; cap = 0;
        CK(cudaMallocHost(&p, n * sizeof(T)));
        cap = n;
    }
    void reserve_growing(size_t n) { if (n > cap) reserve(n + n / 4 + 64); }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// ---------------------------------------------------------------------------------------------------------------
// specialised sweep kernels (ncb_spec.*): generated per program, compiled with nvcc into a cubin cache, loaded through
// the driver API (resolved with dlopen, so the library keeps linking against the runtime only)
// ---------------------------------------------------------------------------------------------------------------
struct DriverApi {
    void* lib = nullptr;
    int (*ModuleLoadData)(void**, const void*) = nullptr;
    int (*ModuleGetFunction)(void**, void*, const char*) = nullptr;
    int (*FuncSetAttribute)(void*, int, int) = nullptr;
    int (*LaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
    int (*ModuleUnload)(void*) = nullptr;
    int (*Occupancy)(int*, void*, int, size_t) = nullptr;
    bool load() {
        if (lib) return true;
        lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        ModuleLoadData = (decltype(ModuleLoadData))dlsym(lib, "cuModuleLoadData");
        ModuleGetFunction = (decltype(ModuleGetFunction))dlsym(lib, "cuModuleGetFunction");
        FuncSetAttribute = (decltype(FuncSetAttribute))dlsym(lib, "cuFuncSetAttribute");
        LaunchKernel = (decltype(LaunchKernel))dlsym(lib, "cuLaunchKernel");
        ModuleUnload = (decltype(ModuleUnload))dlsym(lib, "cuModuleUnload");
        Occupancy = (decltype(Occupancy))dlsym(lib, "cuOccupancyMaxActiveBlocksPerMultiprocessor");
        return ModuleLoadData && ModuleGetFunction && FuncSetAttribute && LaunchKernel && ModuleUnload && Occupancy;
    }
};
static DriverApi g_drv;

// NVRTC (resolved with dlopen as well): the generated kernels are compiled IN PROCESS -- no nvcc binary, no header files on
// disk (the device headers are embedded in the library, embed_headers.py), no child processes.
extern "C" {
extern const int ncb_embedded_count;
extern const char* const ncb_embedded_names[];
extern const char* const ncb_embedded_sources[];
}
struct NvrtcApi {
    void* lib = nullptr;
    bool tried = false;
    int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(void*, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(void*, size_t*) = nullptr;
    int (*GetCUBIN)(void*, char*) = nullptr;
    int (*GetProgramLogSize)(void*, size_t*) = nullptr;
    int (*GetProgramLog)(void*, char*) = nullptr;
    int (*DestroyProgram)(void**) = nullptr;
    bool load() {
        if (tried) return lib != nullptr;
        tried = true;
        std::vector<std::string> names;
        if (const char* e = getenv("NCB_NVRTC_LIB")) names.push_back(e);
        for (const char* n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.13"}) names.push_back(n);
        for (const std::string& n : names) {
            lib = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) return false;
        CreateProgram = (decltype(CreateProgram))dlsym(lib, "nvrtcCreateProgram");
        CompileProgram = (decltype(CompileProgram))dlsym(lib, "nvrtcCompileProgram");
        GetCUBINSize = (decltype(GetCUBINSize))dlsym(lib, "nvrtcGetCUBINSize");
        GetCUBIN = (decltype(GetCUBIN))dlsym(lib, "nvrtcGetCUBIN");
        GetProgramLogSize = (decltype(GetProgramLogSize))dlsym(lib, "nvrtcGetProgramLogSize");
        GetProgramLog = (decltype(GetProgramLog))dlsym(lib, "nvrtcGetProgramLog");
        DestroyProgram = (decltype(DestroyProgram))dlsym(lib, "nvrtcDestroyProgram");
        if (!(CreateProgram && CompileProgram && GetCUBINSize && GetCUBIN && GetProgramLogSize && GetProgramLog && DestroyProgram)) {
            dlclose(lib); lib = nullptr;
        }
        return lib != nullptr;
    }
    // source -> cubin for sm_100a; returns false (with the compiler's log) on failure.  Thread safe per call.
    bool compile(const std::string& src, const std::string& name, std::string& cubin, std::string& log) {
        void* prog = nullptr;
        if (CreateProgram(&prog, src.c_str(), name.c_str(), ncb_embedded_count, ncb_embedded_sources, ncb_embedded_names) != 0) { log = "nvrtcCreateProgram failed"; return false; }
        const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
        const int rc = CompileProgram(prog, 3, opts);
        size_t n = 0;
        if (GetProgramLogSize(prog, &n) == 0 && n > 1) { log.resize(n); GetProgramLog(prog, &log[0]); }
        bool ok = rc == 0;
        if (ok) {
            size_t cn = 0;
            ok = GetCUBINSize(prog, &cn) == 0 && cn > 0;
            if (ok) { cubin.resize(cn); ok = GetCUBIN(prog, &cubin[0]) == 0; }
        }
        DestroyProgram(&prog);
        return ok;
    }
};
static NvrtcApi g_nvrtc;

static std::string build_stamp() {
    uint64_t h = 1469598103934665603ull;
    for (const char* c = __DATE__ " " __TIME__; *c; ++c) { h ^= (unsigned char)*c; h *= 1099511628211ull; }
    char b[20];
    std::snprintf(b, sizeof b, "%08x", (unsigned)(h ^ (h >> 32)));
    return b;
}
static std::string this_library_dir() {
    Dl_info info;
    if (dladdr((void*)&this_library_dir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        const size_t k = p.find_last_of('/');
        return k == std::string::npos ? std::string(".") : p.substr(0, k);
    }
    return ".";
}

struct Engine {
    Program P;
    int device = -1;
    bool uploaded = false;
    int sm_count = 148;
    size_t smem_optin = 0;
    DevProgram d{};
    // the program's tables live in ONE device allocation (d_prog), filled by one copy from a pinned image: a parameter sweep
    // replaces the program once per circuit, and a dozen small synchronous copies cost more than its kernels
    template <class T> struct View { T* p = nullptr; };
    View<Op> d_ops; View<Pass> d_passes; View<Segment> d_segs; View<Instr> d_instrs; View<Channel> d_chans;
    View<double> d_pool; View<int32_t> d_noisy; View<int32_t> d_orig; View<unsigned char> d_blobs;
    View<double> d_thresh; View<int32_t> d_seg_pool; View<ResCkpt> d_ckpt_for;       // tables of the generated resident kernels
    DevBuf<unsigned char> d_prog;
    PinnedBuf<unsigned char> h_prog;
    DevBuf<cplx> d_states;
    DevBuf<double> d_trunk;
    DevBuf<double> d_red, d_norm, d_invnorm, d_weight, d_probs, d_partial, d_uniforms;
    bool overlap_streams = true;
    bool independent = false;          // ncb_options.independent_trajectories: no shared no-jump prefix
    DevBuf<Decision> d_dec;
    DevBuf<unsigned long long> d_counters;
    DevBuf<Work> d_works[2];
    DevBuf<DecideItem> d_decides[2];
    PinnedBuf<Work> h_works[2];
    PinnedBuf<DecideItem> h_decides[2];
    // generated event kernels decide in the kernel (last CTA of a state): per work the event's uniform, per slot a ticket
    DevBuf<double> d_work_u[2];
    PinnedBuf<double> h_work_u[2];
    DevBuf<int> d_tickets;
    const double* cur_work_u = nullptr;
    cudaEvent_t buf_free[2] = {nullptr, nullptr};
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    // second stream: a segment's jump-event chain (first pieces, decide, rounds) runs beside the direct launch of the
    // trajectories without an event in that segment (disjoint slots), joined at every segment boundary
    cudaStream_t s2 = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    bool tile_configured = false, resident_configured = false;
    int tile_ctas = 148;

    ~Engine() {
        d_prog.release(); h_prog.release(); h_cnt.release(); h_probs.release(); d_states.release();
        if (s_own) cudaStreamDestroy(s_own); d_red.release(); d_norm.release(); d_invnorm.release(); d_weight.release(); d_trunk.release(); d_probs.release();
        d_partial.release(); d_uniforms.release(); d_dec.release(); d_counters.release(); d_tickets.release();
        d_nitems.release(); d_ckpt.release(); d_sort_tmp.release();
        for (int i = 0; i < 2; ++i) { d_item_key[i].release(); d_item_val[i].release(); }
        for (int i = 0; i < 2; ++i) {
            d_works[i].release(); d_decides[i].release(); h_works[i].release(); h_decides[i].release();
            d_work_u[i].release(); h_work_u[i].release();
            if (buf_free[i]) cudaEventDestroy(buf_free[i]);
        }
        for (void* m : spec_modules) if (g_drv.ModuleUnload) g_drv.ModuleUnload(m);
        if (s2) cudaStreamDestroy(s2);
        if (ev_a) cudaEventDestroy(ev_a);
        if (ev_b) cudaEventDestroy(ev_b);
        if (t0) cudaEventDestroy(t0);
        if (t1) cudaEventDestroy(t1);
        if (upload_done) cudaEventDestroy(upload_done);
    }

    bool resident() const { return P.nbits <= P.K && P.generics.empty(); }     // (a 5..8-bit unitary is a launch of its own)
    size_t tile_smem() const { return ((size_t)1 << P.K) * sizeof(cplx); }
    size_t resident_smem() const {
        return tile_smem() + ((size_t)1 << P.nbits) * sizeof(double) + (((size_t)P.G + 31) / 32 + 1) * sizeof(unsigned);
    }

    bool device_ready = false;       // device chosen, streams and events created (once); `uploaded`: the program's tables are on it
    cudaEvent_t upload_done = nullptr;
    cudaStream_t upload_stream_used = nullptr;
    void ensure_device(cudaStream_t upload_stream = nullptr) {
        init_device();
        upload_program(upload_stream);
    }
    void init_device() {
        if (device_ready) {
            CK(cudaSetDevice(device));                           // the caller may have switched devices since
        } else {
            int count = 0;
            cudaError_t e = cudaGetDeviceCount(&count);
            if (e != cudaSuccess || count == 0)
                throw CudaError(std::string("no usable CUDA device (") + cudaGetErrorString(e) + "): this engine has no CPU fallback");
            if (device >= 0) CK(cudaSetDevice(device));
            CK(cudaGetDevice(&device));
            cudaDeviceProp prop;
            CK(cudaGetDeviceProperties(&prop, device));
            sm_count = prop.multiProcessorCount;
            smem_optin = prop.sharedMemPerBlockOptin;
            for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&buf_free[i], cudaEventDisableTiming));
            CK(cudaEventCreate(&t0));
            CK(cudaEventCreate(&t1));
            int lo_pri = 0, hi_pri = 0;
            CK(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
            CK(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, hi_pri));
            CK(cudaEventCreateWithFlags(&ev_a, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ev_b, cudaEventDisableTiming));
            if (const char* e2 = getenv("NCB_DIRECT_CHUNK")) direct_chunk = atoi(e2);
            if (const char* e2 = getenv("NCB_EV_CHUNK")) ev_chunk = atoi(e2);
            if (const char* e2 = getenv("NCB_EV_PULL")) ev_pull = atoi(e2) != 0;
            if (const char* e2 = getenv("NCB_OVERLAP")) overlap_streams = atoi(e2) != 0;
            CK(cudaStreamCreateWithFlags(&s_own, cudaStreamNonBlocking));
            d_counters.reserve(4);
            device_ready = true;
        }
    }
    void upload_program(cudaStream_t upload_stream) {
        if (uploaded) return;
        {
            // image: every table at a 256-byte aligned offset
            struct Part { const void* src; size_t bytes; void** dst; };
            std::vector<double> th;
            std::vector<int32_t> sp, ck;
            res_tables = false;
            if (specialize && resident() && spec_resident_supported(P)) {
                th.resize(2 * std::max<size_t>(P.noisy_instrs.size(), 1));
                for (size_t a = 0; a < P.noisy_instrs.size(); ++a) {
                    const Channel& ch = P.chans[P.instrs[P.noisy_instrs[a]].channel];
                    const double* lo = P.pool.data() + ch.bounds;
                    const double* hi = lo + (ch.count - 1);
                    th[2 * a] = ch.base == 0 ? -1.0 : hi[ch.base - 1] + EVENT_MARGIN;             // (plan_events' fast test)
                    th[2 * a + 1] = ch.base == ch.count - 1 ? 2.0 : lo[ch.base] - EVENT_MARGIN;
                }
                const ResidentLayout lay = spec_resident_layout(P);
                for (size_t k = 0; k < lay.pool_src.size(); ++k) { sp.push_back(lay.pool_src[k]); sp.push_back(lay.pool_len[k]); }
                ck = spec_resident_ckpts(P);
                res_tables = true;
            }
            const Part parts[] = {
                {P.ops.data(), P.ops.size() * sizeof(Op), (void**)&d_ops.p}, {P.passes.data(), P.passes.size() * sizeof(Pass), (void**)&d_passes.p},
                {P.segs.data(), P.segs.size() * sizeof(Segment), (void**)&d_segs.p}, {P.instrs.data(), P.instrs.size() * sizeof(Instr), (void**)&d_instrs.p},
                {P.chans.data(), P.chans.size() * sizeof(Channel), (void**)&d_chans.p}, {P.pool.data(), P.pool.size() * sizeof(double), (void**)&d_pool.p},
                {P.noisy_instrs.data(), P.noisy_instrs.size() * sizeof(int32_t), (void**)&d_noisy.p}, {P.orig.data(), P.orig.size() * sizeof(int32_t), (void**)&d_orig.p},
                {P.blobs.data(), P.blobs.size(), (void**)&d_blobs.p}, {th.data(), th.size() * sizeof(double), (void**)&d_thresh.p},
                {sp.data(), sp.size() * sizeof(int32_t), (void**)&d_seg_pool.p}, {ck.data(), ck.size() * sizeof(int32_t), (void**)&d_ckpt_for.p}};
            size_t total = 0;
            for (const Part& pt : parts) total += (pt.bytes + 255) & ~(size_t)255;
            total = std::max<size_t>(total, 256);
            h_prog.reserve_growing(total);
            d_prog.reserve_growing(total);
            size_t at = 0;
            for (const Part& pt : parts) {
                if (pt.bytes) std::memcpy(h_prog.p + at, pt.src, pt.bytes);
                *pt.dst = d_prog.p + at;
                at += (pt.bytes + 255) & ~(size_t)255;
            }
            CK(cudaMemcpyAsync(d_prog.p, h_prog.p, total, cudaMemcpyHostToDevice, upload_stream));
            // (the pinned image is rewritten by the next ncb_program_update at the earliest, after this program's run has been
            //  waited for; runs on other streams are ordered behind the copy by the event)
            if (!upload_done) CK(cudaEventCreateWithFlags(&upload_done, cudaEventDisableTiming));
            CK(cudaEventRecord(upload_done, upload_stream));
            upload_stream_used = upload_stream;
        }
        d.ops = d_ops.p; d.passes = d_passes.p; d.segs = d_segs.p; d.instrs = d_instrs.p; d.chans = d_chans.p;
        d.pool = d_pool.p; d.noisy = d_noisy.p; d.orig = d_orig.p; d.n_noisy = (int)P.noisy_instrs.size(); d.G = P.G; d.nbits = P.nbits;
        d.K = P.K; d.R = P.R; d.nseg = (int)P.segs.size(); d.npass = (int)P.passes.size();
        const size_t need = resident() ? resident_smem() : tile_launch_smem();
        if (need + 6 * 1024 > smem_optin)
            throw std::runtime_error("tile does not fit shared memory: need " + std::to_string(need) + " B, device offers " + std::to_string(smem_optin));
        uploaded = true;
    }
    // A new circuit in the same engine object (ncb_program_update): device buffers, streams and -- when the generated code
    // does not change (same structure, coefficients through the constant bank) -- the loaded kernels are kept.
    std::string spec_key_loaded;
    uint64_t spec_hash_loaded = 0;
    void replace_program(Program&& np) {
        if (pending) throw std::runtime_error("ncb_program_update: a run of this program is in flight (ncb_mcwf_collect first)");
        if (np.K != P.K || np.R != P.R || np.nbits != P.nbits || np.mode != P.mode)
            throw std::runtime_error("ncb_program_update: the new circuit compiles to another geometry (qubits / tile / register bits): create a new program");
        P = std::move(np);
        uploaded = false;
        tile_configured = false;
        direct_ctas = 0;
        bool same = spec_tried && !spec_key_loaded.empty();
        if (same && spec_structure_hash(P, spec_opt) != spec_hash_loaded) {
            same = spec_key(P, spec_opt) == spec_key_loaded;                     // (the hash covers padding bytes: a mismatch is not proof)
            if (getenv("NCB_VERBOSE")) fprintf(stderr, "[ncb] program update: structure hash differs, generated source %s\n", same ? "is the same" : "differs");
            if (same) spec_hash_loaded = spec_structure_hash(P, spec_opt);
        }
        if (spec_tried && !same) {
            for (void* m : spec_modules) if (g_drv.ModuleUnload) g_drv.ModuleUnload(m);
            spec_modules.clear(); spec_fn.clear(); spec_ev_fn.clear();
            res_fn = res_trunk_fn = nullptr;
            spec_tried = false;
            spec_key_loaded.clear();
        }
    }

    // ---- kernel dispatch: three instantiations --------------------------------------------------------------
    //   <3, 0, 512, 1>           R = 3, no dense two-qubit operators (Heron: cz/rzz are diagonal)
    //   <3, DENSE2, 256, 1>      R = 3 with dense 4x4 operators (Eagle ecr, cx, swap, non-diagonal Kraus bases); K <= 11
    //   <4, DENSE2|DENSE4, 256, 1>  R = 4: 3-4 bit dense operators (density-matrix superoperators, 3-4 qubit unitaries)
    int variant() const {
        if (P.R == 4) return (!P.has_dense2 && !P.has_dense4) ? 5 : 2;
        if (P.R == 3 && !P.has_dense4) return P.has_dense2 ? 1 : 0;
        throw std::runtime_error("unsupported reg_bits (3, or 4 for 3-4 bit dense operators)");
    }
    template <class Kern>
    void configure(Kern k) {
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin - 8 * 1024));
        CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    // ---- tile launch configuration (decided once per program) -------------------------------------------------
    int tile_nbuf = 2, tile_blob_cap = MAX_BLOB_BYTES, tile_variant = 0;
    size_t tile_launch_smem() const {
        return (size_t)tile_blob_cap + (size_t)tile_nbuf * tile_smem() + (size_t)JB_TABLE_PASSES * sizeof(unsigned short) * ((size_t)1 << (P.K - P.R));
    }
    template <class Kern>
    int occupancy(Kern k, int threads, size_t smem) {
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, smem));
        return occ;
    }
    // lean kernel variants (R = 3, no dense two-qubit operators): 0: <512,1> 128 regs; 3: <512,2> 64 regs; 4: <256,3> 85 regs
    void configure_tile() {
        if (tile_configured) return;
        int max_blob = 0;
        for (size_t sidx = 0; sidx < P.segs.size(); ++sidx)
            max_blob = std::max(max_blob, reinterpret_cast<const SegBlob*>(P.blobs.data() + P.blob_off[sidx])->bytes);
        tile_blob_cap = (max_blob + 1023) & ~1023;
        tile_variant = variant();
        // lean programs: two 64-register CTAs per SM, one tile buffer each (measured best on B200: the second CTA
        // hides the first one's load latency better than a second buffer does); the register-hungry variants run
        // one CTA per SM and double-buffer instead
        if (tile_variant == 0) { tile_variant = 3; tile_nbuf = 1; }
        if (tile_variant == 5) tile_nbuf = 1;
        if (const char* e = getenv("NCB_TILE_VARIANT")) {
            const int v = atoi(e);
            if ((tile_variant == 0 || tile_variant == 3) && (v == 0 || v == 3 || v == 4)) { tile_variant = v; tile_nbuf = v == 3 ? 1 : 2; }
        }
        if (const char* e = getenv("NCB_NBUF")) tile_nbuf = atoi(e) == 1 ? 1 : 2;
        const int threads = 1 << (P.K - P.R);
        if (tile_variant == 4 && threads > 256) tile_variant = 0;
        configure(tile_kernel<3, 0, 512, 1>);
        configure(tile_kernel<3, 0, 512, 2>);
        configure(tile_kernel<3, 0, 256, 3>);
        configure(tile_kernel<3, FEAT_DENSE2, 256, 1>);
        configure(tile_kernel<4, FEAT_DENSE2 | FEAT_DENSE4, 256, 1>);
        configure(tile_kernel<4, 0, 256, 2>);
        const size_t smem = tile_launch_smem();
        int occ = 1;
        switch (tile_variant) {
        case 0: occ = occupancy(tile_kernel<3, 0, 512, 1>, threads, smem); break;
        case 3: occ = occupancy(tile_kernel<3, 0, 512, 2>, threads, smem); break;
        case 4: occ = occupancy(tile_kernel<3, 0, 256, 3>, threads, smem); break;
        case 1: occ = occupancy(tile_kernel<3, FEAT_DENSE2, 256, 1>, threads, smem); break;
        case 5: occ = occupancy(tile_kernel<4, 0, 256, 2>, threads, smem); break;
        default: occ = occupancy(tile_kernel<4, FEAT_DENSE2 | FEAT_DENSE4, 256, 1>, threads, smem); break;
        }
        if (occ < 1) throw std::runtime_error("tile kernel does not fit on an SM");
        if (const char* e = getenv("NCB_CTAS_PER_SM")) occ = std::max(1, std::min(occ, atoi(e)));
        tile_ctas = sm_count * occ;
        if (getenv("NCB_VERBOSE")) fprintf(stderr, "[ncb] tile kernel variant %d, K=%d, %d threads, nbuf=%d, blob cap %d B, smem %zu B, %d CTAs/SM\n",
                                           tile_variant, P.K, threads, tile_nbuf, tile_blob_cap, smem, occ);
        tile_configured = true;
    }
    void launch_tile(int nworks, cudaStream_t s, const Work* works, int seg) {
        const int tile_shift = P.nbits > P.K ? P.nbits - P.K : 0;
        const int threads = 1 << (P.K - P.R);
        const long long total = (long long)nworks << tile_shift;
        configure_tile();
        const size_t smem = tile_launch_smem();
        const int grid = (int)std::min<long long>(total, tile_ctas);
        if (grid < 1) return;
        const unsigned char* blob = d_blobs.p + P.blob_off[seg];
        const int blob_bytes = reinterpret_cast<const SegBlob*>(P.blobs.data() + P.blob_off[seg])->bytes;
#define NCB_TILE_ARGS d, P.segconst[seg], blob, blob_bytes, tile_blob_cap, tile_nbuf, works, nworks, d_states.p, tile_shift, d_red.p, d_norm.p, d_dec.p
        switch (tile_variant) {
        case 0: tile_kernel<3, 0, 512, 1><<<grid, threads, smem, s>>>(NCB_TILE_ARGS); break;
        case 3: tile_kernel<3, 0, 512, 2><<<grid, threads, smem, s>>>(NCB_TILE_ARGS); break;
        case 4: tile_kernel<3, 0, 256, 3><<<grid, threads, smem, s>>>(NCB_TILE_ARGS); break;
        case 1: tile_kernel<3, FEAT_DENSE2, 256, 1><<<grid, threads, smem, s>>>(NCB_TILE_ARGS); break;
        case 5: tile_kernel<4, 0, 256, 2><<<grid, threads, smem, s>>>(NCB_TILE_ARGS); break;
        default: tile_kernel<4, FEAT_DENSE2 | FEAT_DENSE4, 256, 1><<<grid, threads, smem, s>>>(NCB_TILE_ARGS); break;
        }
#undef NCB_TILE_ARGS
        CK(cudaGetLastError());
    }
    // direct kernel (works that run a whole segment without an event; lean programs only)
    int ev_chunk = 0;
    bool ev_pull = true;          // generated event kernels keep merged groups whole around an in-line jump they do not touch (WF_PULL)
    int direct_ctas = 0, direct_chunk = 8;       // tiles per direct-kernel CTA: CTA turnover lets the (higher-priority) event chain in
                                                 // (measured on the 20-qubit workload: 8 -> 2612, 16 -> 2557, 32 -> 2531, persistent -> slower)
    size_t direct_smem() const {
        return (size_t)tile_blob_cap + tile_smem() + (size_t)SC_MAX_PASSES * sizeof(unsigned short) * ((size_t)1 << (P.K - P.R));
    }
    void launch_direct(int nworks, cudaStream_t s, const Work* works, int seg) {
        configure_tile();
        const int tile_shift = P.nbits - P.K, threads = 1 << (P.K - P.R);
        if (!direct_ctas) {
            configure(direct_kernel<512, 2>);
            int occ = occupancy(direct_kernel<512, 2>, threads, direct_smem());
            if (occ < 1) throw std::runtime_error("direct kernel does not fit on an SM");
            if (const char* e = getenv("NCB_CTAS_PER_SM")) occ = std::max(1, std::min(occ, atoi(e)));
            direct_ctas = sm_count * occ;
            if (getenv("NCB_VERBOSE")) fprintf(stderr, "[ncb] direct kernel: %d threads, smem %zu B, %d CTAs/SM\n", threads, direct_smem(), occ);
        }
        const long long total = (long long)nworks << tile_shift;
        long long want = direct_ctas;
        if (direct_chunk > 0) want = std::max<long long>(want, (total + direct_chunk - 1) / direct_chunk);   // CTA turnover lets the event chain in
        const int grid = (int)std::min<long long>(total, want);
        if (grid < 1) return;
        const unsigned char* blob = d_blobs.p + P.blob_off[seg];
        const int blob_bytes = reinterpret_cast<const SegBlob*>(P.blobs.data() + P.blob_off[seg])->bytes;
        direct_kernel<512, 2><<<grid, threads, direct_smem(), s>>>(d, P.segconst[seg], blob, blob_bytes, tile_blob_cap, works, nworks, d_states.p, tile_shift);
        CK(cudaGetLastError());
    }
    // ---- specialised kernels -----------------------------------------------------------------------------------
    bool specialize = false, spec_tried = false;
    int specialize_mode = 0;          // ncb_options.specialize: 1 literal coefficients, 2 coefficients through the constant bank
    SpecOptions spec_opt;
    int spec_pool_cap = 0;
    std::vector<void*> spec_modules;
    struct SpecFn { void* fn = nullptr; int ctas = 0; };       // CUfunction + resident CTAs per SM (nullptr: interpreter)
    std::vector<SpecFn> spec_fn, spec_ev_fn;                   // per segment: whole-segment kernel, event-piece kernel
    static constexpr int SPEC_PER_PART = 4;                  // segments per translation unit
    void read_spec_options() {
        spec_opt = SpecOptions();
        if (P.K >= 12 || P.R == 4) spec_opt.nbuf = 1;   // K = 12: two 64 KB buffers would leave one CTA per SM; R = 4 (128-thread
                                                        // CTAs): four single-buffered CTAs per SM beat three double-buffered ones
        if (const char* e = getenv("NCB_SPEC_NBUF")) spec_opt.nbuf = atoi(e) == 1 ? 1 : 2;
        if (const char* e = getenv("NCB_SPEC_LOAD")) spec_opt.load = std::min(std::max(atoi(e), 0), 2);
        if (const char* e = getenv("NCB_SPEC_EVENTS")) spec_opt.events = atoi(e) != 0;
        if (const char* e = getenv("NCB_SPEC_CTAS")) spec_opt.ctas = std::max(0, atoi(e));
        if (const char* e = getenv("NCB_SPEC_FACTOR")) spec_opt.factor = atoi(e) != 0;
        if (const char* e = getenv("NCB_RES_FAST")) spec_opt.resident_fast = atoi(e) != 0;
        if (const char* e = getenv("NCB_EV_FAST")) spec_opt.ev_fast = atoi(e) != 0;
        spec_opt.literal = specialize_mode == 2 ? 0 : 1;
        if (const char* e = getenv("NCB_SPEC_LITERAL")) spec_opt.literal = atoi(e) != 0;
    }
    // Cubin cache: $NCB_SPEC_CACHE, else <library dir>/_speccache, else (read-only installation) ~/.cache/noisycircuits_b200/
    // speccache, else the temporary directory.  One sub-directory per (generated source, library build); the least recently
    // used ones are removed beyond NCB_SPEC_CACHE_MAX (default 64) entries.
    static std::string spec_cache_root(const std::string& libdir) {
        std::vector<std::string> cands;
        if (const char* e = getenv("NCB_SPEC_CACHE")) cands.push_back(e);
        else {
            cands.push_back(libdir + "/_speccache");
            if (const char* h = getenv("HOME")) cands.push_back(std::string(h) + "/.cache/noisycircuits_b200/speccache");
            const char* t = getenv("TMPDIR");
            cands.push_back(std::string(t ? t : "/tmp") + "/noisycircuits_b200_speccache");
        }
        for (const std::string& c : cands) {
            // mkdir -p
            for (size_t k = 1; k <= c.size(); ++k)
                if (k == c.size() || c[k] == '/') mkdir(c.substr(0, k).c_str(), 0755);
            if (access(c.c_str(), W_OK | X_OK) == 0) return c;
        }
        return "";
    }
    static void spec_cache_evict(const std::string& cache) {
        size_t keep = 64;
        if (const char* e = getenv("NCB_SPEC_CACHE_MAX")) keep = (size_t)std::max(1, atoi(e));
        std::vector<std::pair<time_t, std::string>> dirs;
        if (DIR* d = opendir(cache.c_str())) {
            while (dirent* e = readdir(d)) {
                if (e->d_name[0] == '.') continue;
                const std::string p = cache + "/" + e->d_name;
                struct stat st;
                if (stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) dirs.push_back({st.st_mtime, p});
            }
            closedir(d);
        }
        if (dirs.size() < keep) return;
        std::sort(dirs.begin(), dirs.end());
        for (size_t i = 0; i + keep <= dirs.size(); ++i) {          // oldest first; leaves keep - 1 (a new one is about to be added)
            if (DIR* d = opendir(dirs[i].second.c_str())) {
                while (dirent* e = readdir(d))
                    if (e->d_name[0] != '.') unlink((dirs[i].second + "/" + e->d_name).c_str());
                closedir(d);
            }
            rmdir(dirs[i].second.c_str());
        }
    }
    // Generates and compiles (NVRTC, in parallel, into the cubin cache) the specialised kernels of this program; host only.
    // Returns the cache directory of this program ("" on failure: the interpreter stays in charge, and says so in
    // ncb_run_params.out_spec_launches).
    std::string spec_build() {
        const bool verbose = getenv("NCB_VERBOSE") != nullptr;
        int nsupported = 0;
        for (int sgi = 0; sgi < (int)P.segs.size(); ++sgi) nsupported += spec_supported(P, sgi);
        const bool res = resident() && spec_resident_supported(P);       // one more part: the resident kernels
        if (!nsupported && !res) return "";
        read_spec_options();
        if (res && verbose) {
            const ResidentLayout lay = spec_resident_layout(P, spec_opt);
            fprintf(stderr, "[ncb] generated resident kernels: %d B of staged tables, %zu B shared memory, accumulator in %s memory, %d CTAs/SM planned\n",
                    lay.pool_bytes, lay.smem, lay.acc_global ? "global" : "shared", lay.ctas);
        }
        const std::string libdir = this_library_dir();
        const std::string cache = spec_cache_root(libdir);
        if (cache.empty()) { if (verbose) fprintf(stderr, "[ncb] specialised kernels: no writable cache directory\n"); return ""; }
        // (the cubins depend on the kernel headers too: a rebuilt library starts a new cache generation)
        const std::string dir = cache + "/" + spec_key(P, spec_opt) + "-" + build_stamp();
        struct stat st_dir;
        if (stat(dir.c_str(), &st_dir) != 0) {
            spec_cache_evict(cache);
            mkdir(dir.c_str(), 0755);
        } else {
            utimes(dir.c_str(), nullptr);            // most recently used
        }
        const int nseg = (int)P.segs.size();
        const int per = SPEC_PER_PART;
        const int nparts = nsupported ? (nseg + per - 1) / per : 0;
        auto exists = [](const std::string& f) { struct stat st; return stat(f.c_str(), &st) == 0 && st.st_size > 0; };
        auto part_name = [&](int k) { return k == nparts ? std::string("resident") : k == nparts + 1 ? std::string("restrunk") : "part" + std::to_string(k); };
        std::vector<int> todo;
        for (int k = 0; k < nparts + (res ? 2 : 0); ++k)
            if (!exists(dir + "/" + part_name(k) + ".cubin")) todo.push_back(k);
        if (!todo.empty()) {
            // compiler: NVRTC in process (default); NCB_SPEC_COMPILER=nvcc runs the nvcc binary instead (needs the toolkit
            // and this library's csrc/ headers on disk)
            const char* want = getenv("NCB_SPEC_COMPILER");
            bool use_nvrtc = !(want && std::string(want) == "nvcc") && g_nvrtc.load();
            const std::string nvcc = getenv("NCB_NVCC") ? getenv("NCB_NVCC") : "/usr/local/cuda/bin/nvcc";
            if (!use_nvrtc && !exists(nvcc)) {
                if (verbose) fprintf(stderr, "[ncb] specialised kernels: neither libnvrtc nor %s is usable\n", nvcc.c_str());
                return "";
            }
            // several processes (one per GPU) build the same cache at once: each starts at another part and skips what a
            // neighbour has finished meanwhile
            {
                int rot = 0;
                if (const char* e = getenv("LOCAL_RANK")) rot = atoi(e);
                int ranks = 1;
                if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
                std::rotate(todo.begin(), todo.begin() + (size_t)((long long)rot * (long long)todo.size() / ranks) % todo.size(), todo.end());
            }
            std::atomic<int> next{0}, failed{0};
            auto worker = [&]() {
                for (;;) {
                    const int i = next.fetch_add(1);
                    if (i >= (int)todo.size()) break;
                    const int k = todo[i];
                    if (exists(dir + "/" + part_name(k) + ".cubin")) continue;
                    // several processes (one per GPU) may build the same cache at once: private temporaries, atomic renames
                    const std::string base = dir + "/" + part_name(k), mine = base + "." + std::to_string((long)getpid());
                    const std::string src = k >= nparts ? spec_resident_source(P, spec_opt, k == nparts ? 1 : 2) : spec_source(P, k * per, std::min(nseg, (k + 1) * per), spec_opt);
                    { std::ofstream f(mine + ".cu"); f << src; }
                    bool ok;
                    if (use_nvrtc) {
                        std::string cubin, log;
                        ok = g_nvrtc.compile(src, part_name(k) + ".cu", cubin, log);
                        { std::ofstream f(mine + ".log"); f << log; }
                        if (ok) { std::ofstream f(mine + ".tmp", std::ios::binary); f.write(cubin.data(), (std::streamsize)cubin.size()); ok = (bool)f; }
                        if (ok) ok = rename((mine + ".tmp").c_str(), (base + ".cubin").c_str()) == 0;
                        rename((mine + ".cu").c_str(), (base + ".cu").c_str());
                        rename((mine + ".log").c_str(), (base + ".log").c_str());
                    } else {
                        const std::string cmd = nvcc + " -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I" + libdir + "/csrc -Xptxas -v -cubin -o " +
                                                mine + ".tmp " + mine + ".cu > " + mine + ".log 2>&1 && mv " + mine + ".tmp " + base + ".cubin && mv " +
                                                mine + ".cu " + base + ".cu && mv " + mine + ".log " + base + ".log";
                        ok = system(cmd.c_str()) == 0;
                    }
                    if (!ok) failed.fetch_add(1);
                }
            };
            const int nthreads = std::max(1, std::min<int>({(int)todo.size(), plan_threads(), 16}));
            std::vector<std::thread> pool;
            for (int i = 0; i < nthreads; ++i) pool.emplace_back(worker);
            for (auto& t : pool) t.join();
            if (verbose) fprintf(stderr, "[ncb] specialised kernels: compiled %zu parts with %s on %d threads (%d failed) into %s\n", todo.size(),
                                 use_nvrtc ? "NVRTC" : "nvcc", nthreads, failed.load(), dir.c_str());
            if (failed.load()) return "";
        }
        return dir;
    }
    size_t spec_smem(bool ev) const { return (size_t)spec_pool_cap + (size_t)(ev ? 1 : (spec_opt.nbuf == 1 ? 1 : 2)) * tile_smem(); }
    void ensure_spec() {
        if (spec_tried) return;
        spec_tried = true;
        spec_fn.assign(P.segs.size(), SpecFn());
        spec_ev_fn.assign(P.segs.size(), SpecFn());
        const bool verbose = getenv("NCB_VERBOSE") != nullptr;
        if (!g_drv.load()) { if (verbose) fprintf(stderr, "[ncb] specialised kernels: libcuda.so.1 not usable\n"); return; }
        const std::string dir = spec_build();
        if (dir.empty()) return;
        spec_key_loaded = spec_key(P, spec_opt);
        spec_hash_loaded = spec_structure_hash(P, spec_opt);
        const int nseg = (int)P.segs.size(), per = SPEC_PER_PART;
        int nparts = (nseg + per - 1) / per;
        if (resident()) {
            nparts = 0;
            load_resident_module(dir, verbose);
        }
        spec_pool_cap = 0;
        for (int sgi = 0; sgi < nseg; ++sgi) {
            int off, bytes;
            spec_pool(P, sgi, &off, &bytes);
            spec_pool_cap = std::max(spec_pool_cap, (bytes + 1023) & ~1023);
        }
        const int threads = 1 << (P.K - P.R);
        int loaded = 0, loaded_ev = 0;
        auto get = [&](void* mod, const std::string& name, bool ev) {
            SpecFn f;
            void* fn = nullptr;
            if (g_drv.ModuleGetFunction(&fn, mod, name.c_str()) != 0) return f;
            if (g_drv.FuncSetAttribute(fn, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, (int)smem_optin - 8 * 1024) != 0) return f;
            g_drv.FuncSetAttribute(fn, 9 /* CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT */, 100);
            int occ = 0;
            if (g_drv.Occupancy(&occ, fn, threads, spec_smem(ev)) != 0 || occ < 1) return f;
            if (const char* e = getenv("NCB_CTAS_PER_SM")) occ = std::max(1, std::min(occ, atoi(e)));
            f.fn = fn; f.ctas = occ;
            return f;
        };
        for (int k = 0; k < nparts; ++k) {
            std::ifstream f(dir + "/part" + std::to_string(k) + ".cubin", std::ios::binary);
            std::stringstream ss;
            ss << f.rdbuf();
            const std::string image = ss.str();
            void* mod = nullptr;
            if (image.empty() || g_drv.ModuleLoadData(&mod, image.data()) != 0) continue;
            spec_modules.push_back(mod);
            for (int sgi = k * per; sgi < std::min(nseg, (k + 1) * per); ++sgi) {
                if (spec_supported(P, sgi)) { spec_fn[sgi] = get(mod, "ncb_seg_" + std::to_string(sgi), false); loaded += spec_fn[sgi].fn != nullptr; }
                if (spec_opt.events && spec_ev_supported(P, sgi)) {
                    spec_ev_fn[sgi] = get(mod, "ncb_seg_" + std::to_string(sgi) + "_ev", true);
                    loaded_ev += spec_ev_fn[sgi].fn != nullptr;
                }
            }
        }
        if (verbose) fprintf(stderr, "[ncb] specialised kernels: %d whole-segment and %d event kernels of %d segments loaded (load mode %d, nbuf %d, %d / %d CTAs per SM)\n",
                             loaded, loaded_ev, nseg, spec_opt.load, spec_opt.nbuf, nseg && spec_fn[0].fn ? spec_fn[0].ctas : 0, nseg && spec_ev_fn[0].fn ? spec_ev_fn[0].ctas : 0);
    }
    // ---- generated resident kernels (ncb_res_trunk / ncb_res, see ncb_spec.cpp) ----------------------------------------
    void* res_fn = nullptr;
    void* res_trunk_fn = nullptr;
    ResidentLayout res_lay;
    int res_ctas = 0;
    DevBuf<int32_t> d_nitems, d_item_val[2];
    DevBuf<uint32_t> d_item_key[2];
    DevBuf<unsigned char> d_sort_tmp;
    DevBuf<cplx> d_ckpt;
    bool res_tables = false;
    void load_resident_module(const std::string& dir, bool verbose) {
        res_fn = res_trunk_fn = nullptr;
        if (!spec_resident_supported(P)) return;
        res_lay = spec_resident_layout(P, spec_opt);
        auto load = [&](const char* file, const char* name) -> void* {
            std::ifstream f(dir + "/" + file, std::ios::binary);
            std::stringstream ss;
            ss << f.rdbuf();
            const std::string image = ss.str();
            void *mod = nullptr, *fn = nullptr;
            if (image.empty() || g_drv.ModuleLoadData(&mod, image.data()) != 0) return nullptr;
            spec_modules.push_back(mod);
            return g_drv.ModuleGetFunction(&fn, mod, name) == 0 ? fn : nullptr;
        };
        void* fr = load("resident.cubin", "ncb_res");
        void* ft = load("restrunk.cubin", "ncb_res_trunk");
        if (!fr || !ft) return;
        for (void* fn : {fr, ft}) {
            if (g_drv.FuncSetAttribute(fn, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, (int)smem_optin - 8 * 1024) != 0) return;
            g_drv.FuncSetAttribute(fn, 9 /* CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT */, 100);
        }
        int occ = 0;
        if (g_drv.Occupancy(&occ, fr, 1 << (P.K - P.R), res_lay.smem) != 0 || occ < 1) return;
        if (const char* e = getenv("NCB_CTAS_PER_SM")) occ = std::max(1, std::min(occ, atoi(e)));
        res_ctas = occ;
        res_fn = fr; res_trunk_fn = ft;
        if (verbose) fprintf(stderr, "[ncb] generated resident kernels loaded: %d threads, %zu B shared memory (%d B of tables), accumulator in %s memory, %d CTAs/SM\n",
                             1 << (P.K - P.R), res_lay.smem, res_lay.pool_bytes, res_lay.acc_global ? "global" : "shared", occ);
    }
    // checkpoints of the trunk (the tables of the generated resident kernels travel with the program image, ensure_device)
    void ensure_res_tables() {
        if (!res_tables) throw std::runtime_error("internal: resident tables missing");
        d_ckpt.reserve(((size_t)P.passes.size() + 1) << P.K);
        d_nitems.reserve(1);
    }

    // returns true when the step ran in a kernel generated for the program
    bool launch_step(const LaunchStep& st, cudaStream_t s, const Work* works) {
        if (specialize) {
            ensure_spec();
            const SpecFn f = st.seg < (int)spec_fn.size() ? (st.direct ? spec_fn[st.seg] : spec_ev_fn[st.seg]) : SpecFn();
            if (f.fn) {
                const int tile_shift = P.nbits - P.K, threads = 1 << (P.K - P.R);
                const long long total = (long long)st.count << tile_shift;
                long long want = (long long)sm_count * f.ctas;
                if (st.direct && direct_chunk > 0) want = std::max<long long>(want, (total + direct_chunk - 1) / direct_chunk);   // CTA turnover lets the event chain in
                // event launches are small (a few trajectories): with ev_chunk > 0 a CTA takes at least that many tiles, so a
                // launch occupies fewer CTA slots for longer -- its latency hides behind the segment's whole-segment launch
                if (!st.direct && ev_chunk > 0) want = std::min<long long>(want, std::max<long long>(sm_count, (total + ev_chunk - 1) / ev_chunk));
                const int grid = (int)std::min<long long>(total, want);
                if (grid < 1) return true;
                int pool_off, pool_bytes;
                spec_pool(P, st.seg, &pool_off, &pool_bytes);
                const unsigned char* pool_g = d_blobs.p + P.blob_off[st.seg] + pool_off;
                int pool_cap = spec_pool_cap, nworks = st.count, tshift = tile_shift;
                const Work* w = works + st.begin;
                cplx* states = d_states.p;
                int rc;
                if (st.direct) {
                    double* nrm = d_norm.p;
                    void* args[] = {(void*)P.segconst[st.seg].coef, (void*)&pool_g, &pool_bytes, &pool_cap, (void*)&w, &nworks, (void*)&states, &tshift, (void*)&nrm};
                    rc = g_drv.LaunchKernel(f.fn, (unsigned)grid, 1, 1, (unsigned)threads, 1, 1, (unsigned)spec_smem(false), (void*)s, args, nullptr);
                } else {
                    double *red = d_red.p, *nrm = d_norm.p;
                    Decision* dec = d_dec.p;
                    const double* wu = cur_work_u ? cur_work_u + st.begin : nullptr;
                    int* tk = cur_work_u ? d_tickets.p : nullptr;          // (null: a decide launch follows, NCB_FUSED_DECIDE=0)
                    int pull_on = ev_pull ? 1 : 0;
                    void* args[] = {(void*)P.segconst[st.seg].coef, (void*)&d, (void*)&pool_g, &pool_bytes, &pool_cap, (void*)&w, &nworks, (void*)&states, &tshift,
                                    (void*)&red, (void*)&nrm, (void*)&dec, (void*)&wu, (void*)&tk, &pull_on};
                    rc = g_drv.LaunchKernel(f.fn, (unsigned)grid, 1, 1, (unsigned)threads, 1, 1, (unsigned)spec_smem(true), (void*)s, args, nullptr);
                }
                if (rc != 0) throw std::runtime_error("cuLaunchKernel of a specialised kernel failed (" + std::to_string(rc) + ")");
                return true;
            }
        }
        if (st.direct && P.R == 3) launch_direct(st.count, s, works + st.begin, st.seg);      // (the interpreting direct kernel is R = 3)
        else launch_tile(st.count, s, works + st.begin, st.seg);
        return false;
    }
    // a segment that is one 5..8-bit unitary
    void launch_generic(const LaunchStep& st, cudaStream_t s, const Work* works) {
        const GenericOp* go = nullptr;
        for (const GenericOp& g : P.generics) if (g.seg == st.seg) go = &g;
        if (!go) throw std::runtime_error("internal: generic launch on an ordinary segment");
        GenericArgs ga;
        ga.nb = go->nb;
        for (int i = 0; i < 8; ++i) ga.bits[i] = go->bits[i];
        const long long groups = ((long long)st.count << (P.nbits - go->nb)) >> (8 - go->nb);      // block iterations
        const int grid = (int)std::max<long long>(1, std::min<long long>(groups, (long long)sm_count * 8));
        generic_unitary_kernel<<<grid, 256, 0, s>>>(works + st.begin, st.count, d_states.p, P.nbits, ga, reinterpret_cast<const cplx*>(d_pool.p + go->mat));
        CK(cudaGetLastError());
    }
    void launch_resident(int grid, cudaStream_t s, const ResidentArgs& a) {
        const int threads = 1 << (P.K - P.R);
        if (!resident_configured) {
            configure(resident_kernel<3, 0, 512, 1>);
            configure(resident_kernel<3, FEAT_DENSE2, 256, 1>);
            configure(resident_kernel<4, FEAT_DENSE2 | FEAT_DENSE4, 256, 1>);
            resident_configured = true;
        }
        switch (variant()) {
        case 0: resident_kernel<3, 0, 512, 1><<<grid, threads, resident_smem(), s>>>(d, a); break;
        case 1: resident_kernel<3, FEAT_DENSE2, 256, 1><<<grid, threads, resident_smem(), s>>>(d, a); break;
        default: resident_kernel<4, FEAT_DENSE2 | FEAT_DENSE4, 256, 1><<<grid, threads, resident_smem(), s>>>(d, a); break;
        }
        CK(cudaGetLastError());
    }
    int resident_ctas_per_sm() const {
        const size_t per = resident_smem() + 8 * 1024;
        int by_smem = (int)std::max<size_t>(1, (size_t)227 * 1024 / per);
        int by_threads = std::max(1, 1536 / (1 << (P.K - P.R)));
        return std::max(1, std::min({by_smem, by_threads, 8}));
    }

    // uniform(t_local, instr) for the host planner
    struct UniformSource {
        int rng; uint64_t seed; long long traj_begin; const double* table; int G;
        std::vector<double> mt;         // [count][num_draws]
        int num_draws = 0;
        const int32_t* draw_pos = nullptr;
        double get(long long tl, int i) const {
            if (rng == NCB_RNG_PHILOX) return philox_uniform(seed, (uint64_t)(traj_begin + tl), (uint32_t)i);
            if (rng == NCB_RNG_EXPLICIT) return table[tl * G + i];
            return mt[(size_t)tl * num_draws + draw_pos[i]];
        }
    };

    static double mt_canonical(std::mt19937_64& g) {
        // what std::discrete_distribution draws: std::generate_canonical<double, 53>(g)
        return std::generate_canonical<double, 53>(g);
    }

    void fill_mt(UniformSource& us, long long begin, long long count) {
        us.num_draws = P.num_draws; us.draw_pos = P.draw_pos.data();
        us.mt.resize((size_t)count * std::max(P.num_draws, 1));
#pragma omp parallel for schedule(static)
        for (long long t = 0; t < count; ++t) {
            std::mt19937_64 g((unsigned int)(us.seed + (uint64_t)(us.traj_begin + begin + t)));   // Simulator.cpp:124: cuint seed
            for (int k = 0; k < P.num_draws; ++k) us.mt[(size_t)t * P.num_draws + k] = mt_canonical(g);
        }
    }

    // dense per-instruction uniform table [count][G] (resident path with MT / explicit rng)
    void dense_table(const UniformSource& us, long long begin, long long count, std::vector<double>& out) {
        out.assign((size_t)count * P.G, 0.5);
#pragma omp parallel for schedule(static)
        for (long long t = 0; t < count; ++t)
            for (int32_t pos : P.noisy_instrs) out[(size_t)t * P.G + pos] = us.get(begin + t, P.orig[pos]);   // by position
    }

    // host threads of the event planner: launchers like torchrun export OMP_NUM_THREADS=1, which would serialise the
    // planning of a batch (tens of ms at the start of every call); take this rank's share of the cores instead
    static int plan_threads() {
        if (const char* e = getenv("NCB_PLAN_THREADS")) return std::max(1, atoi(e));
        int hw = (int)std::thread::hardware_concurrency();
        if (hw < 1) hw = 1;
        int ranks = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
        return std::max(1, std::min(16, hw / ranks));
    }
    int auto_batch(long long count) const {
        // enough trajectories per launch that (a) the persistent grid runs many tiles per CTA and (b) the small extra
        // launches of the jump-event rounds (a few slots each) are amortised; capped by memory (8 GB of state)
        const long long tiles = 1ll << std::max(P.nbits - P.K, 0);
        long long b = (148ll * 1024 + tiles - 1) / tiles;
        b = ((b + 36) / 37) * 37;
        if (const char* e = getenv("NCB_BATCH")) b = std::max(1, atoi(e));
        long long mem_gb = 16;
        if (const char* e = getenv("NCB_BATCH_MEM_GB")) mem_gb = std::max(1, atoi(e));
        const long long mem_cap = std::max<long long>(1, (mem_gb << 30) / ((long long)sizeof(cplx) << P.nbits));
        b = std::min(b, mem_cap);
        return (int)std::max<long long>(1, std::min(b, count));
    }

    // Two-phase run: mcwf_begin enqueues, mcwf_finish waits and delivers (ncb_mcwf_run = both; ncb_mcwf_launch / _collect let
    // a caller keep several programs in flight).  `deferred`: nothing has been waited for yet -- only the generated resident path
    // can do that (its planning happens on the device); every other path has finished its kernels when mcwf_begin returns.
    bool pending = false, deferred = false;
    cudaStream_t s_own = nullptr, s_run = nullptr;     // the program's own stream (two-phase runs with stream = 0); stream of the pending run
    PinnedBuf<unsigned long long> h_cnt;
    PinnedBuf<double> h_probs;
    void mcwf_begin(ncb_run_params* rp, bool may_defer);
    void mcwf_finish(ncb_run_params* rp);
    void mcwf_run(ncb_run_params* rp) { mcwf_begin(rp, false); mcwf_finish(rp); }
    void mcwf_resident(ncb_run_params* rp, cudaStream_t s, bool defer = false);
    void mcwf_tiled(ncb_run_params* rp, cudaStream_t s);
    void run_deterministic(cudaStream_t s);       // one slot, no events (pure / density modes); leaves the state in d_states
};

static void grid_1d(long long n, int* grid, int* block) {
    *block = 256;
    *grid = (int)std::min<long long>((n + 255) / 256, 148 * 16);
    if (*grid < 1) *grid = 1;
}

void Engine::mcwf_resident(ncb_run_params* rp, cudaStream_t s, bool defer) {
    const long long dim = 1ll << P.nbits, count = rp->traj_count;
    const int grid = (int)std::min<long long>(count, (long long)sm_count * resident_ctas_per_sm());
    UniformSource us{rp->rng, rp->seed, rp->traj_begin, rp->uniforms, P.G};
    ResidentArgs a{};
    a.traj_begin = rp->traj_begin; a.traj_count = count; a.seed = rp->seed; a.normalize = P.norm_preserving ? 0 : 1;
    a.uniforms = nullptr;
    if (rp->rng != NCB_RNG_PHILOX) {
        std::vector<double> table;
        if (rp->rng == NCB_RNG_MT19937_REFERENCE) fill_mt(us, 0, count);
        dense_table(us, 0, count, table);
        d_uniforms.reserve(table.size());
        CK(cudaMemcpyAsync(d_uniforms.p, table.data(), table.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        rp->out_h2d_bytes += (int64_t)(table.size() * sizeof(double));
        a.uniforms = d_uniforms.p;
    }
    if (rp->states) { d_states.reserve((size_t)count * dim); a.states_out = d_states.p; }
    CK(cudaMemsetAsync(d_counters.p, 0, 4 * sizeof(unsigned long long), s));
    a.counters = d_counters.p;
    a.share = (!rp->states && !independent) ? 1 : 0;
    if (const char* e = getenv("NCB_SHARE_PREFIX")) a.share = a.share && atoi(e) != 0;
    d_trunk.reserve(dim);
    a.trunk_probs = d_trunk.p;
    int g, b;
    grid_1d(dim, &g, &b);
    bool generated = false;
    if (specialize && !independent) {
        ensure_spec();
        generated = res_fn != nullptr;
    }
    if (generated) {
        // generated kernels: the trunk once (second stream) beside the planning kernels, then the trajectories that leave it
        ensure_res_tables();
        const int nblk = (int)((count + RES_PLAN_THREADS / 32 - 1) / (RES_PLAN_THREADS / 32));
        for (int i = 0; i < 2; ++i) { d_item_key[i].reserve_growing((size_t)count); d_item_val[i].reserve_growing((size_t)count); }
        int key_bits = 1;
        while ((1 << key_bits) <= P.G + 1) ++key_bits;
        size_t tmp_bytes = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_item_key[0].p, d_item_key[1].p, d_item_val[0].p, d_item_val[1].p, (int)count, 0, key_bits, s));
        d_sort_tmp.reserve_growing(tmp_bytes + 16);
        // (one CTA slot of the device stays free: the trunk of the NEXT program of a sweep runs beside this launch)
        const int rgrid = (int)std::max<long long>(1, std::min<long long>(count, (long long)sm_count * res_ctas - 1));
        d_partial.reserve((size_t)rgrid * dim);
        ResArgs ra{};
        ra.traj_begin = rp->traj_begin; ra.seed = rp->seed; ra.uniforms = a.uniforms; ra.thresh = d_thresh.p; ra.blobs = d_blobs.p;
        ra.seg_pool = d_seg_pool.p; ra.item_first = d_item_key[1].p; ra.item_traj = d_item_val[1].p; ra.n_items = d_nitems.p; ra.count = (int)count;
        ra.ckpt_for = d_ckpt_for.p; ra.ckpt = d_ckpt.p;
        ra.partial = d_partial.p; ra.states_out = a.states_out; ra.trunk_probs = d_trunk.p; ra.counters = d_counters.p;
        ra.normalize = a.normalize; ra.acc_global = res_lay.acc_global;
        const unsigned threads = 1u << (P.K - P.R);
        void* args[] = {(void*)&d, (void*)&ra};
        cudaStream_t st = s2;
        CK(cudaEventRecord(ev_a, s));
        CK(cudaStreamWaitEvent(st, ev_a, 0));
        int rc = g_drv.LaunchKernel(res_trunk_fn, 1, 1, 1, threads, 1, 1, (unsigned)res_lay.smem, (void*)st, args, nullptr);
        if (rc != 0) throw std::runtime_error("cuLaunchKernel of the generated trunk kernel failed (" + std::to_string(rc) + ")");
        CK(cudaEventRecord(ev_b, st));
        CK(cudaMemsetAsync(d_nitems.p, 0, sizeof(int32_t), s));
        res_plan_kernel<<<nblk, RES_PLAN_THREADS, 0, s>>>(d_thresh.p, d_noisy.p, d_orig.p, (int)P.noisy_instrs.size(), P.G, rp->seed, rp->traj_begin, (int)count,
                                                          a.uniforms, rp->states ? 1 : 0, d_item_key[0].p, d_item_val[0].p, d_nitems.p);
        CK(cudaGetLastError());
        // stable sort by the first flagged draw: longest remaining work first, unflagged trajectories (key G + 1) last
        tmp_bytes = d_sort_tmp.cap;
        CK(cub::DeviceRadixSort::SortPairs(d_sort_tmp.p, tmp_bytes, d_item_key[0].p, d_item_key[1].p, d_item_val[0].p, d_item_val[1].p, (int)count, 0, key_bits, s));
        CK(cudaStreamWaitEvent(s, ev_b, 0));
        rc = g_drv.LaunchKernel(res_fn, (unsigned)rgrid, 1, 1, threads, 1, 1, (unsigned)res_lay.smem, (void*)s, args, nullptr);
        if (rc != 0) throw std::runtime_error("cuLaunchKernel of the generated resident kernel failed (" + std::to_string(rc) + ")");
        sum_partials_wide_kernel<<<(unsigned)((dim + 31) / 32), 256, 0, s>>>(d_probs.p, d_partial.p, dim, rgrid, d_trunk.p, d_counters.p + 2);
        CK(cudaGetLastError());
        rp->out_spec_launches += 2;
        rp->out_launches += 3;                   // (+ the radix sort's)
    } else {
        d_partial.reserve((size_t)grid * dim);
        a.partial_probs = d_partial.p;
        launch_resident(grid, s, a);
        sum_partials_kernel<<<g, b, 0, s>>>(d_probs.p, d_partial.p, dim, grid, a.share ? d_trunk.p : nullptr, d_counters.p + 2);
        CK(cudaGetLastError());
    }
    h_cnt.reserve(4);
    unsigned long long* cnt = h_cnt.p;
    CK(cudaMemcpyAsync(cnt, d_counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    rp->out_d2h_bytes += (int64_t)(4 * sizeof(unsigned long long));
    rp->out_tile_launches += 1;
    rp->out_launches += 2;
    if (defer && generated && !rp->states) { deferred = true; return; }       // (mcwf_finish adds the event counts)
    if (rp->states) {
        CK(cudaMemcpyAsync(rp->states, d_states.p, (size_t)count * dim * sizeof(cplx), cudaMemcpyDeviceToHost, s));
        rp->out_d2h_bytes += (int64_t)((size_t)count * dim * sizeof(cplx));
    }
    CK(cudaStreamSynchronize(s));
    if (rp->states && !(P.global_phase[0] == 1.0 && P.global_phase[1] == 0.0))
        for (size_t j = 0; j < (size_t)count * dim; ++j) {
            const double x = rp->states[2 * j], y = rp->states[2 * j + 1];
            rp->states[2 * j] = x * P.global_phase[0] - y * P.global_phase[1];
            rp->states[2 * j + 1] = x * P.global_phase[1] + y * P.global_phase[0];
        }
    rp->out_events += (int64_t)cnt[0];
    rp->out_hard_events += (int64_t)cnt[1];
}

void Engine::mcwf_tiled(ncb_run_params* rp, cudaStream_t s) {
    const long long dim = 1ll << P.nbits, count = rp->traj_count;
    const int ntiles = 1 << (P.nbits > P.K ? P.nbits - P.K : 0);     // (a small state with a 5..8-bit unitary runs as one tile)
    const int B = rp->batch > 0 ? (int)std::min<long long>(rp->batch, count) : auto_batch(count);
    // trajectories are one and the same evolution until their first jump: computed once per batch (ncb_schedule.h)
    bool share = !rp->states && !independent;
    if (const char* e = getenv("NCB_SHARE_PREFIX")) share = share && atoi(e) != 0;
    const int SL = B + 2;                       // state slots: B trajectories + the two trunk slots
    d_states.reserve((size_t)SL * dim);
    d_red.reserve((size_t)SL * ntiles * RED_VALUES);
    d_norm.reserve((size_t)SL * ntiles);
    d_invnorm.reserve(SL);
    d_weight.reserve(SL);
    d_dec.reserve(SL);
    UniformSource us{rp->rng, rp->seed, rp->traj_begin, rp->uniforms, P.G};
    std::vector<std::vector<Event>> events;
    BatchPlan plan;
    std::vector<cplx> host_states;
    int buf = 0;
    const bool want_norm = !P.norm_preserving;
    // generated kernels: the whole-segment kernel of the last segment accumulates the squared norm itself
    bool direct_norm = false;
    if (specialize && want_norm) {
        ensure_spec();
        direct_norm = !spec_fn.empty() && spec_fn.back().fn != nullptr;
    }
    CK(cudaMemsetAsync(d_red.p, 0, (size_t)SL * ntiles * RED_VALUES * sizeof(double), s));
    d_tickets.reserve(SL);
    CK(cudaMemsetAsync(d_tickets.p, 0, (size_t)SL * sizeof(int), s));
    // NCB_FUSED_DECIDE=1: the generated event kernels decide the branch themselves (last CTA of a state) instead of a decide
    // launch between the pieces.  Off by default: measured 3 % slower (the chain's latency is hidden behind the segment's
    // whole-segment launch anyway, while the deciding CTA's sum over 512 partials sits in the event launch's tail).
    bool fused_decide_on = false;
    if (const char* e = getenv("NCB_FUSED_DECIDE")) fused_decide_on = atoi(e) != 0;
    fused_decide_on = fused_decide_on && specialize && (ensure_spec(), !spec_ev_fn.empty());
    for (long long b0 = 0; b0 < count; b0 += B) {
        const int nb = (int)std::min<long long>(B, count - b0);
        if (rp->rng == NCB_RNG_MT19937_REFERENCE) {
            UniformSource tmp = us;
            fill_mt(tmp, b0, nb);
            us.mt.swap(tmp.mt); us.num_draws = tmp.num_draws; us.draw_pos = tmp.draw_pos;
        }
        events.assign(nb, {});
        const bool mt = rp->rng == NCB_RNG_MT19937_REFERENCE;
#pragma omp parallel for schedule(dynamic, 4) num_threads(plan_threads())
        for (int b = 0; b < nb; ++b) {
            const long long tl = mt ? b : b0 + b;
            plan_events(P, [&](int i) { return us.get(tl, i); }, events[b]);
        }
        schedule_batch(P, events, true, want_norm, plan, share, true, direct_norm);
        for (int b = 0; b < nb; ++b) {
            rp->out_events += (int64_t)events[b].size();
            for (const Event& e : events[b]) rp->out_hard_events += e.kmin != e.kmax;
        }
        rp->out_sweeps += plan.sweeps;
        // upload the batch's work lists (double-buffered so planning overlaps the previous batch's kernels)
        CK(cudaEventSynchronize(buf_free[buf]));
        h_works[buf].reserve_growing(plan.works.size() + 1); d_works[buf].reserve_growing(plan.works.size() + 1);
        h_decides[buf].reserve_growing(plan.decides.size() + 1); d_decides[buf].reserve_growing(plan.decides.size() + 1);
        std::memcpy(h_works[buf].p, plan.works.data(), plan.works.size() * sizeof(Work));
        CK(cudaMemcpyAsync(d_works[buf].p, h_works[buf].p, plan.works.size() * sizeof(Work), cudaMemcpyHostToDevice, s));
        rp->out_h2d_bytes += (int64_t)(plan.works.size() * sizeof(Work) + plan.decides.size() * sizeof(DecideItem));
        if (!plan.decides.empty()) {
            std::memcpy(h_decides[buf].p, plan.decides.data(), plan.decides.size() * sizeof(DecideItem));
            CK(cudaMemcpyAsync(d_decides[buf].p, h_decides[buf].p, plan.decides.size() * sizeof(DecideItem), cudaMemcpyHostToDevice, s));
        }
        cur_work_u = nullptr;
        if (fused_decide_on) {
            // the uniform of the event each reducing work ends in (the generated event kernels decide in the kernel)
            h_work_u[buf].reserve_growing(plan.works.size() + 1); d_work_u[buf].reserve_growing(plan.works.size() + 1);
            for (size_t i = 0; i < plan.works.size(); ++i) {
                const Work& w = plan.works[i];
                double u = 0.0;
                if ((w.flags & WF_REDUCE) && w.slot < nb)
                    for (const Event& e : events[w.slot])
                        if (e.instr == w.instr_hi - 1) { u = e.u; break; }
                h_work_u[buf].p[i] = u;
            }
            CK(cudaMemcpyAsync(d_work_u[buf].p, h_work_u[buf].p, plan.works.size() * sizeof(double), cudaMemcpyHostToDevice, s));
            rp->out_h2d_bytes += (int64_t)(plan.works.size() * sizeof(double));
            cur_work_u = d_work_u[buf].p;
        }
        const bool overlap = overlap_streams && share;
        auto join = [&]() {          // both streams wait for each other
            CK(cudaEventRecord(ev_a, s));
            CK(cudaEventRecord(ev_b, s2));
            CK(cudaStreamWaitEvent(s, ev_b, 0));
            CK(cudaStreamWaitEvent(s2, ev_a, 0));
        };
        int cur_seg = -1;
        if (overlap) join();         // the uploads above and the previous batch's tail
        for (const LaunchStep& st : plan.steps) {
            if (overlap && st.seg != cur_seg && cur_seg >= 0) join();
            cur_seg = st.seg;
            cudaStream_t q = overlap && !(st.kind == 0 && st.direct) && st.kind != 2 ? s2 : s;
            if (st.kind == 0) { rp->out_spec_launches += launch_step(st, q, d_works[buf].p) ? 1 : 0; ++rp->out_tile_launches; }
            else if (st.kind == 2) { launch_generic(st, q, d_works[buf].p); ++rp->out_tile_launches; }
            else {
                // (the generated event kernel of this segment has decided already: last CTA of every reducing state)
                if (fused_decide_on && st.seg < (int)spec_ev_fn.size() && spec_ev_fn[st.seg].fn) continue;
                decide_kernel<<<st.count, 256, 0, q>>>(d, d_decides[buf].p + st.begin, st.count, d_red.p, ntiles, d_dec.p);
                CK(cudaGetLastError());
            }
            ++rp->out_launches;
        }
        if (overlap) join();
        CK(cudaEventRecord(buf_free[buf], s));
        CK(cudaMemcpyAsync(d_weight.p, plan.weight.data(), plan.weight.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        rp->out_h2d_bytes += (int64_t)(plan.weight.size() * sizeof(double));
        if (want_norm) {
            norm_finalize_kernel<<<(plan.nslots + 7) / 8, 256, 0, s>>>(d_norm.p, ntiles, plan.nslots, d_weight.p, d_invnorm.p);
            CK(cudaGetLastError());
        }
        int g, bl;
        grid_1d(dim, &g, &bl);
        accumulate_kernel<<<g, bl, 0, s>>>(d_probs.p, d_states.p, P.nbits, plan.nslots, want_norm ? d_invnorm.p : d_weight.p);
        CK(cudaGetLastError());
        rp->out_launches += want_norm ? 2 : 1;
        if (rp->states) {
            host_states.resize((size_t)nb * dim);
            CK(cudaMemcpyAsync(host_states.data(), d_states.p, (size_t)nb * dim * sizeof(cplx), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            rp->out_d2h_bytes += (int64_t)((size_t)nb * dim * sizeof(cplx));
            for (int b = 0; b < nb; ++b) {
                double n2 = 0;
                for (long long j = 0; j < dim; ++j) n2 += cnorm2(host_states[(size_t)b * dim + j]);
                const double f = 1.0 / std::sqrt(n2);
                const double gx = P.global_phase[0] * f, gy = P.global_phase[1] * f;
                for (long long j = 0; j < dim; ++j) {
                    const cplx a = host_states[(size_t)b * dim + j];
                    rp->states[2 * ((size_t)(b0 + b) * dim + j)] = a.x * gx - a.y * gy;
                    rp->states[2 * ((size_t)(b0 + b) * dim + j) + 1] = a.x * gy + a.y * gx;
                }
            }
        }
        buf ^= 1;
    }
    CK(cudaStreamSynchronize(s));
}

void Engine::mcwf_begin(ncb_run_params* rp, bool may_defer) {
    if (pending) throw std::runtime_error("a run of this program is still in flight: call ncb_mcwf_collect first");
    if (P.mode != MODE_MCWF) throw std::runtime_error("program was not compiled in MCWF mode");
    if (rp->traj_count < 0) throw std::runtime_error("traj_count must be >= 0");
    if (rp->rng == NCB_RNG_EXPLICIT && !rp->uniforms) throw std::runtime_error("rng = EXPLICIT needs uniforms");
    if (rp->rng == NCB_RNG_MT19937_REFERENCE) {
        // the reference's stream is consumed in PROGRAM order and its branch decisions see the state of that order
        // (Simulator.cpp:76-94): a re-ordered program would only be the reference's trajectory in distribution
        for (int i = 0; i < P.G; ++i)
            if (P.orig[i] != i) throw std::runtime_error("rng = MT19937_REFERENCE needs a program compiled with strict_order = 1");
    }
    if (!rp->probs) throw std::runtime_error("probs must not be NULL");
    cudaStream_t s = (cudaStream_t)rp->stream;
    init_device();
    if (may_defer && !s) s = s_own;      // two-phase runs without a caller's stream use the program's own: several programs overlap on the device
    ensure_device(s);
    if (upload_done && s != upload_stream_used) CK(cudaStreamWaitEvent(s, upload_done, 0));
    const long long dim = 1ll << P.nbits;
    d_probs.reserve(dim);
    CK(cudaMemsetAsync(d_probs.p, 0, dim * sizeof(double), s));
    rp->out_sweeps = rp->out_events = rp->out_hard_events = rp->out_launches = 0;
    rp->out_tile_launches = rp->out_h2d_bytes = rp->out_d2h_bytes = rp->out_spec_launches = 0;
    rp->out_kernel_ms = 0.0;
    deferred = false;
    CK(cudaEventRecord(t0, s));
    if (rp->traj_count > 0) {
        if (resident()) {
            // chunk so the optional tables / states stay bounded
            const long long chunk = rp->states || rp->rng != NCB_RNG_PHILOX ? std::max<long long>(1, (1ll << 28) / std::max<long long>(dim * 2, (long long)P.G))
                                                                            : std::min<long long>(rp->traj_count, 1ll << 24);
            const bool one = chunk >= rp->traj_count;
            ncb_run_params sub = *rp;
            for (long long b0 = 0; b0 < rp->traj_count; b0 += chunk) {
                sub.traj_begin = rp->traj_begin + b0;
                sub.traj_count = std::min(chunk, rp->traj_count - b0);
                sub.uniforms = rp->uniforms ? rp->uniforms + b0 * P.G : nullptr;
                sub.states = rp->states ? rp->states + 2 * b0 * dim : nullptr;
                mcwf_resident(&sub, s, may_defer && one && rp->rng == NCB_RNG_PHILOX);
            }
            rp->out_sweeps = sub.out_sweeps; rp->out_events = sub.out_events; rp->out_hard_events = sub.out_hard_events;
            rp->out_launches = sub.out_launches; rp->out_tile_launches = sub.out_tile_launches;
            rp->out_h2d_bytes = sub.out_h2d_bytes; rp->out_d2h_bytes = sub.out_d2h_bytes; rp->out_spec_launches = sub.out_spec_launches;
        } else {
            mcwf_tiled(rp, s);
        }
    }
    CK(cudaEventRecord(t1, s));
    // deliver (the host copy is issued here and read in mcwf_finish)
    if (rp->probs_on_device) {
        if (rp->accumulate) {
            // probs += d_probs
            int g, b;
            grid_1d(dim, &g, &b);
            sum_partials_kernel<<<g, b, 0, s>>>(rp->probs, d_probs.p, dim, 1);
            CK(cudaGetLastError());
        } else {
            CK(cudaMemcpyAsync(rp->probs, d_probs.p, dim * sizeof(double), cudaMemcpyDeviceToDevice, s));
        }
    } else {
        h_probs.reserve((size_t)dim);
        CK(cudaMemcpyAsync(h_probs.p, d_probs.p, dim * sizeof(double), cudaMemcpyDeviceToHost, s));
        rp->out_d2h_bytes += (int64_t)(dim * sizeof(double));
    }
    s_run = s;
    pending = true;
}

void Engine::mcwf_finish(ncb_run_params* rp) {
    if (!pending) throw std::runtime_error("no run of this program is in flight");
    pending = false;
    CK(cudaSetDevice(device));
    CK(cudaStreamSynchronize(s_run));
    {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, t0, t1));
        rp->out_kernel_ms = ms;
    }
    if (deferred) {
        rp->out_events += (int64_t)h_cnt.p[0];
        rp->out_hard_events += (int64_t)h_cnt.p[1];
        deferred = false;
    }
    if (!rp->probs_on_device) {
        const long long dim = 1ll << P.nbits;
        if (rp->accumulate) for (long long j = 0; j < dim; ++j) rp->probs[j] += h_probs.p[j];
        else std::memcpy(rp->probs, h_probs.p, dim * sizeof(double));
    }
}

void Engine::run_deterministic(cudaStream_t s) {
    ensure_device(s);
    const long long dim = 1ll << P.nbits;
    d_states.reserve(dim);
    if (resident()) {
        ResidentArgs a{};
        a.traj_begin = 0; a.traj_count = 1; a.seed = 0; a.uniforms = nullptr; a.partial_probs = nullptr;
        a.states_out = d_states.p; a.normalize = 0; a.counters = nullptr;
        launch_resident(1, s, a);
    } else {
        d_red.reserve(RED_VALUES); d_norm.reserve((size_t)1 << (P.nbits > P.K ? P.nbits - P.K : 0)); d_dec.reserve(1);
        std::vector<std::vector<Event>> events(1);
        BatchPlan plan;
        schedule_batch(P, events, true, false, plan);
        d_works[0].reserve(plan.works.size() + 1);
        CK(cudaMemcpyAsync(d_works[0].p, plan.works.data(), plan.works.size() * sizeof(Work), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        for (const LaunchStep& st : plan.steps) {
            if (st.kind == 2) launch_generic(st, s, d_works[0].p);
            else launch_step(st, s, d_works[0].p);
        }
    }
}

}  // namespace ncb

// ==================================================================================================================
// C ABI
// ==================================================================================================================
using namespace ncb;

struct ncb_program {
    Engine eng;
};

static int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}
#define NCB_TRY try {
#define NCB_CATCH                                                           \
    }                                                                       \
    catch (const CudaError& e) { return fail(NCB_E_CUDA, e.what()); }       \
    catch (const std::bad_alloc&) { return fail(NCB_E_NOMEM, "out of host memory"); } \
    catch (const std::exception& e) { return fail(NCB_E_INVALID, e.what()); }

extern "C" {

const char* ncb_last_error(void) { return g_error.c_str(); }
int ncb_version(void) { return NCB_VERSION; }

static CircuitView view_of(const ncb_circuit* c) {
    CircuitView v;
    v.num_qubits = c->num_qubits; v.num_instructions = c->num_instructions;
    v.opcode = c->opcode; v.q0 = c->q0; v.q1 = c->q1; v.theta = c->theta; v.channel = c->channel;
    v.unitary_offset = c->unitary_offset; v.unitary_nq = c->unitary_nq; v.unitary_qubits = c->unitary_qubits;
    v.unitary_pool = c->unitary_pool; v.num_channels = c->num_channels; v.channel_offset = c->channel_offset;
    v.channel_count = c->channel_count; v.channel_dim = c->channel_dim; v.kraus_pool = c->kraus_pool;
    return v;
}

int ncb_program_create(const ncb_circuit* circuit, int mode, const ncb_options* options, ncb_program** out) {
    if (!circuit || !out) return fail(NCB_E_INVALID, "null argument");
    NCB_TRY
    CompileOptions opt;
    int device = -1;
    // default register bits: 3; 4 when the program carries 3-4 bit dense operators (density-matrix superoperators,
    // 3-4 qubit unitaries)
    opt.reg_bits = mode == NCB_MODE_DENSITY ? 4 : 3;
    bool need4 = false;
    for (int g = 0; g < circuit->num_instructions; ++g)
        if (circuit->opcode[g] == NCB_UNITARY && circuit->unitary_nq && circuit->unitary_nq[g] >= 3) { opt.reg_bits = 4; need4 = true; }
    // generated kernels (ncb_options.specialize) of lean tiled MCWF programs run best with R = 4 (16 amplitudes per thread,
    // 128-thread CTAs, a quarter fewer passes: measured +10 % on the 20-qubit workload); the interpreting kernels with R = 3
    int spec_mode = options ? options->specialize : 0;
    if (const char* e = getenv("NCB_SPECIALIZE")) spec_mode = atoi(e);
    const bool reg_explicit = (options && options->reg_bits > 0) || getenv("NCB_REG_BITS") != nullptr;
    const bool auto4 = mode == NCB_MODE_MCWF && (spec_mode == 1 || spec_mode == 2) && !reg_explicit && !need4;
    if (auto4) opt.reg_bits = 4;
    if (options) {
        if (options->tile_bits > 0) { opt.tile_bits = options->tile_bits; opt.resident_bits = options->tile_bits; }
        if (options->reg_bits > 0) opt.reg_bits = options->reg_bits;
        if (options->low_bits >= 0) opt.low_bits = options->low_bits;
        opt.fuse = options->fuse;
        opt.reorder = options->strict_order ? 0 : 1;
        device = options->device;
    }
    if (const char* e = getenv("NCB_TILE_BITS")) { opt.tile_bits = atoi(e); opt.resident_bits = opt.tile_bits; }
    if (const char* e = getenv("NCB_REG_BITS")) opt.reg_bits = atoi(e);
    if (const char* e = getenv("NCB_LOW_BITS")) opt.low_bits = atoi(e);
    if (const char* e = getenv("NCB_FUSE")) opt.fuse = atoi(e);
    if (const char* e = getenv("NCB_REORDER")) opt.reorder = atoi(e);
    if (const char* e = getenv("NCB_DECOMPOSE")) opt.decompose = atoi(e);
    if (const char* e = getenv("NCB_DIRECT")) opt.direct = atoi(e);
    if (const char* e = getenv("NCB_FUSE_ROT")) opt.fuse_rotations = atoi(e);
    if (const char* e = getenv("NCB_SPLIT_U")) opt.split_generic = atoi(e);
    if (const char* e = getenv("NCB_PASS_ORDER")) opt.pass_order = atoi(e);
    if (auto4 && mode == NCB_MODE_MCWF && circuit->num_qubits <= std::max(std::min(opt.tile_bits, 12), std::min(opt.resident_bits, MAX_TILE_BITS - 1)))
        opt.reg_bits = 3;                       // the state fits one CTA: R = 3 (see below), known before compiling
    if (opt.reg_bits != 3 && opt.reg_bits != 4) return fail(NCB_E_INVALID, "reg_bits must be 3 or 4");
    opt.tile_bits = std::min(opt.tile_bits, 12);   // <= 256 threads (R = 4) / 512 threads (R = 3)
    ncb_program* p = new ncb_program();
    try {
        compile_program(view_of(circuit), mode, opt, p->eng.P);
        if (auto4) {
            // R = 4 pays for rotation / diagonal programs (Heron: +10 %); programs full of generic 2x2 operators (Eagle: the
            // one-qubit factors of the decomposed ecr) run 7 % faster with R = 3 (measured, profiles/r02_variants.jsonl)
            // (counted on the fast lists, what the whole-segment kernels run: the one-qubit factors that hand their phases to
            //  a neighbouring table run as rotations there, CompileOptions::split_generic -- with it Eagle is R = 4 too: 19.3 k
            //  against 17.7 k trajectories/s)
            long long rx = 0, u = 0;
            for (const Pass& ps : p->eng.P.passes)
                for (int o = ps.fop_begin; o < ps.fop_end; ++o) { rx += op_kind(p->eng.P.ops[o]) == OPK_RX1; u += op_kind(p->eng.P.ops[o]) == OPK_DENSE1; }
            const bool generic_heavy = 4 * u > rx + u;
            // (states that fit one CTA -- generated resident kernels -- run faster with R = 3: twice the threads per state,
            //  0.90 against 1.21 ms on the 12-qubit QNN circuit)
            if (p->eng.P.has_dense2 || p->eng.P.has_dense4 || p->eng.P.nbits <= p->eng.P.K || generic_heavy) {
                opt.reg_bits = 3;                   // (or not a program the generated kernels cover at all: the default geometry)
                compile_program(view_of(circuit), mode, opt, p->eng.P);
            }
        }
        if (p->eng.P.R == 3 && p->eng.P.has_dense2 && p->eng.P.K > 11) {     // the dense-4x4 kernels run 256 threads
            opt.tile_bits = 11;
            opt.resident_bits = std::min(opt.resident_bits, 11);             // (a 12-bit state then runs as two tiles)
            compile_program(view_of(circuit), mode, opt, p->eng.P);
        }
    } catch (...) { delete p; throw; }
    p->eng.device = device;
    p->eng.independent = options && options->independent_trajectories != 0;
    p->eng.specialize_mode = options ? options->specialize : 0;
    if (const char* e = getenv("NCB_SPECIALIZE")) p->eng.specialize_mode = atoi(e);
    p->eng.specialize = p->eng.specialize_mode == 1 || p->eng.specialize_mode == 2;
    *out = p;
    return NCB_OK;
    NCB_CATCH
}

void ncb_program_destroy(ncb_program* program) { delete program; }

int ncb_program_update(ncb_program* program, const ncb_circuit* circuit) {
    if (!program || !circuit) return fail(NCB_E_INVALID, "null argument");
    NCB_TRY
    Engine& E = program->eng;
    Program np;
    compile_program(view_of(circuit), E.P.mode, E.P.opt, np);
    E.replace_program(std::move(np));
    return NCB_OK;
    NCB_CATCH
}

int64_t ncb_program_describe(const ncb_program* program, char* buf, int64_t len) {
    if (!program) return 0;
    const std::string s = describe_program(program->eng.P);
    if (buf && len > 0) {
        const size_t n = std::min<size_t>(s.size(), (size_t)len - 1);
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int64_t)s.size() + 1;
}

int64_t ncb_program_query(const ncb_program* program, int what) {
    if (!program) return -1;
    const Program& P = program->eng.P;
    switch (what) {
    case NCB_Q_NUM_QUBITS: return P.nq;
    case NCB_Q_STATE_BITS: return P.nbits;
    case NCB_Q_TILE_BITS: return P.K;
    case NCB_Q_REG_BITS: return P.R;
    case NCB_Q_NUM_SEGMENTS: return (int64_t)P.segs.size();
    case NCB_Q_NUM_PASSES: return (int64_t)P.passes.size();
    case NCB_Q_NUM_OPS: return (int64_t)P.ops.size();
    case NCB_Q_NUM_DRAWS: return P.num_draws;
    case NCB_Q_NORM_PRESERVING: return P.norm_preserving;
    case NCB_Q_NUM_INSTRUCTIONS: return P.G;
    case NCB_Q_RESIDENT: return program->eng.resident();
    }
    return -1;
}

// Generates and compiles the specialised kernels of the program into the cubin cache (host only: nvcc, no device).
int ncb_program_spec_build(ncb_program* program) {
    if (!program) return fail(NCB_E_INVALID, "null argument");
    NCB_TRY
    return program->eng.spec_build().empty() ? fail(NCB_E_INVALID, "specialised kernels could not be built (unsupported program or nvcc failure)") : NCB_OK;
    NCB_CATCH
}

// Source of the specialised kernels of segments [seg_begin, seg_end) (ncb_spec.cpp); returns the bytes needed.
int64_t ncb_program_spec_source(const ncb_program* program, int32_t seg_begin, int32_t seg_end, char* buf, int64_t len) {
    if (!program) return -1;
    const Program& P = program->eng.P;
    Engine& E = const_cast<ncb_program*>(program)->eng;
    E.read_spec_options();
    // (a program whose state fits one CTA has one module: the generated resident kernels)
    const std::string s = E.resident() ? spec_resident_source(P, E.spec_opt)
                                       : spec_source(P, std::max(seg_begin, 0), std::min<int>(seg_end, (int)P.segs.size()), E.spec_opt);
    if (buf && len > 0) {
        const size_t n = std::min<size_t>(s.size(), (size_t)len - 1);
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int64_t)s.size() + 1;
}

int ncb_program_order(const ncb_program* program, int32_t* out) {
    if (!program || !out) return fail(NCB_E_INVALID, "null argument");
    const Program& P = program->eng.P;
    for (int i = 0; i < P.G; ++i) out[i] = P.orig[i];
    return NCB_OK;
}

int64_t ncb_program_events(const ncb_program* program, const double* uniforms, int32_t* out, int64_t cap) {
    if (!program || !uniforms) return -1;
    const Program& P = program->eng.P;
    std::vector<Event> ev;
    plan_events(P, [&](int i) { return uniforms[i]; }, ev);
    for (size_t i = 0; i < ev.size() && (int64_t)i < cap; ++i) {
        out[3 * i] = P.orig[ev[i].instr]; out[3 * i + 1] = ev[i].kmin; out[3 * i + 2] = ev[i].kmax;
    }
    return (int64_t)ev.size();
}

int ncb_program_reference_uniforms(const ncb_program* program, uint64_t seed, double* uniforms) {
    if (!program || !uniforms) return fail(NCB_E_INVALID, "null argument");
    const Program& P = program->eng.P;
    std::mt19937_64 g((unsigned int)seed);
    for (int i = 0; i < P.G; ++i) uniforms[i] = P.draw_pos[i] >= 0 ? Engine::mt_canonical(g) : 0.5;
    return NCB_OK;
}

int ncb_mcwf_run(ncb_program* program, ncb_run_params* params) {
    if (!program || !params) return fail(NCB_E_INVALID, "null argument");
    NCB_TRY
    program->eng.mcwf_run(params);
    return NCB_OK;
    NCB_CATCH
}

int ncb_mcwf_launch(ncb_program* program, ncb_run_params* params) {
    if (!program || !params) return fail(NCB_E_INVALID, "null argument");
    NCB_TRY
    program->eng.mcwf_begin(params, true);
    return NCB_OK;
    NCB_CATCH
}

int ncb_mcwf_collect(ncb_program* program, ncb_run_params* params) {
    if (!program || !params) return fail(NCB_E_INVALID, "null argument");
    NCB_TRY
    program->eng.mcwf_finish(params);
    return NCB_OK;
    NCB_CATCH
}

int ncb_pure_run(ncb_program* program, double* state, double* probs, void* stream) {
    if (!program) return fail(NCB_E_INVALID, "null argument");
    NCB_TRY
    Engine& E = program->eng;
    if (E.P.mode != MODE_PURE) throw std::runtime_error("program was not compiled in PURE mode");
    cudaStream_t s = (cudaStream_t)stream;
    E.run_deterministic(s);
    const long long dim = 1ll << E.P.nbits;
    if (state) CK(cudaMemcpyAsync(state, E.d_states.p, dim * sizeof(cplx), cudaMemcpyDeviceToHost, s));
    if (probs) {
        E.d_probs.reserve(dim);
        int g, b;
        grid_1d(dim, &g, &b);
        abs2_kernel<<<g, b, 0, s>>>(E.d_probs.p, E.d_states.p, dim);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(probs, E.d_probs.p, dim * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    if (state && !(E.P.global_phase[0] == 1.0 && E.P.global_phase[1] == 0.0))
        for (long long j = 0; j < dim; ++j) {
            const double x = state[2 * j], y = state[2 * j + 1];
            state[2 * j] = x * E.P.global_phase[0] - y * E.P.global_phase[1];
            state[2 * j + 1] = x * E.P.global_phase[1] + y * E.P.global_phase[0];
        }
    return NCB_OK;
    NCB_CATCH
}

int ncb_density_run(ncb_program* program, double* probs, void* stream) {
    if (!program || !probs) return fail(NCB_E_INVALID, "null argument");
    NCB_TRY
    Engine& E = program->eng;
    if (E.P.mode != MODE_DENSITY) throw std::runtime_error("program was not compiled in DENSITY mode");
    cudaStream_t s = (cudaStream_t)stream;
    E.run_deterministic(s);
    const long long dim = 1ll << E.P.nq;
    E.d_probs.reserve(dim);
    int g, b;
    grid_1d(dim, &g, &b);
    diag_kernel<<<g, b, 0, s>>>(E.d_probs.p, E.d_states.p, E.P.nq);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(probs, E.d_probs.p, dim * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return NCB_OK;
    NCB_CATCH
}

int ncb_apply_measurement_error_device(double* probabilities_device, const double* mats, const int32_t* qubits,
                                       int32_t nq, int32_t num_qubits, void* stream) {
    if (!probabilities_device || !mats || !qubits) return fail(NCB_E_INVALID, "null argument");
    NCB_TRY
    if (num_qubits < 1 || num_qubits > MAX_QUBITS) throw std::runtime_error("num_qubits out of range");
    const long long half = (1ll << num_qubits) >> 1;
    int g, b;
    grid_1d(half, &g, &b);
    for (int k = 0; k < nq; ++k) {
        if (qubits[k] < 0 || qubits[k] >= num_qubits) throw std::runtime_error("measured qubit out of range");
        measurement_error_kernel<<<g, b, 0, (cudaStream_t)stream>>>(probabilities_device, half, qubits[k], mats[4 * k],
                                                                     mats[4 * k + 1], mats[4 * k + 2], mats[4 * k + 3]);
        CK(cudaGetLastError());
    }
    return NCB_OK;
    NCB_CATCH
}

int ncb_execute_tail_device(const double* probabilities_device, int32_t num_qubits, const int32_t* qubits, int32_t nq,
                            const double* mats, double scale, double* out, void* stream) {
    if (!probabilities_device || !qubits || !out) return fail(NCB_E_INVALID, "null argument");
    NCB_TRY
    if (num_qubits < 1 || num_qubits > MAX_QUBITS) throw std::runtime_error("num_qubits out of range");
    if (nq < 1 || nq > num_qubits) throw std::runtime_error("number of measured qubits out of range");
    unsigned keep = 0;
    for (int k = 0; k < nq; ++k) {
        if (qubits[k] < 0 || qubits[k] >= num_qubits) throw std::runtime_error("measured qubit out of range");
        if ((keep >> qubits[k]) & 1u) throw std::runtime_error("measured qubit listed twice");
        keep |= 1u << qubits[k];
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long long nout = 1ll << nq;
    // scratch: kept between calls (cudaMalloc / cudaFree cost up to 100 ms each next to a multi-GB state batch -- measured)
    static std::mutex tail_mutex;
    static DevBuf<double> tail_scratch[2];
    static int tail_device = -1;
    std::lock_guard<std::mutex> lock(tail_mutex);
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev != tail_device) { tail_scratch[0] = DevBuf<double>(); tail_scratch[1] = DevBuf<double>(); tail_device = dev; }   // (per device; the old buffers stay with their device)
    tail_scratch[0].reserve((size_t)nout);
    tail_scratch[1].reserve((size_t)nout);
    double *a = tail_scratch[0].p, *b = tail_scratch[1].p;
    int rc = NCB_OK;
    {
        int g, bl;
        grid_1d(nout, &g, &bl);
        marginal_kernel<<<g, bl, 0, s>>>(a, probabilities_device, num_qubits, keep, nq, scale);
        CK(cudaGetLastError());
        // the marginal is little-endian over the kept qubits in ascending order: qubit q sits at bit rank(q) of it.  (The
        // reference strides by the qubit id itself, MeasurementErrorApplicator.cpp:109-112, which reads out of bounds for
        // any measured set that is not a permutation of 0..nq-1; here every subset works and mats[k] belongs to qubits[k].)
        std::vector<int32_t> ranks(nq);
        for (int k = 0; k < nq; ++k) ranks[k] = __builtin_popcount(keep & ((1u << qubits[k]) - 1u));
        if (mats) rc = ncb_apply_measurement_error_device(a, mats, ranks.data(), nq, nq, stream);
        if (rc == NCB_OK) {
            bit_reverse_kernel<<<g, bl, 0, s>>>(b, a, nq);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(out, b, sizeof(double) * nout, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
        }
    }
    return rc;
    NCB_CATCH
}

int ncb_device_alloc(int64_t bytes, void** out_device_pointer) {
    if (!out_device_pointer || bytes < 0) return fail(NCB_E_INVALID, "bad argument");
    NCB_TRY
    *out_device_pointer = nullptr;
    CK(cudaMalloc(out_device_pointer, (size_t)std::max<int64_t>(bytes, 1)));
    return NCB_OK;
    NCB_CATCH
}

int ncb_device_free(void* device_pointer) {
    NCB_TRY
    if (device_pointer) CK(cudaFree(device_pointer));
    return NCB_OK;
    NCB_CATCH
}

int ncb_apply_measurement_error(double* probabilities, const double* mats, const int32_t* qubits, int32_t nq,
                                int32_t num_qubits, uint32_t num_threads) {
    (void)num_threads;
    if (!probabilities) return fail(NCB_E_INVALID, "null argument");
    NCB_TRY
    if (num_qubits < 1 || num_qubits > MAX_QUBITS) throw std::runtime_error("num_qubits out of range");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) throw CudaError(std::string("no usable CUDA device (") + cudaGetErrorString(e) + ")");
    const size_t bytes = sizeof(double) << num_qubits;
    double* d = nullptr;
    CK(cudaMalloc(&d, bytes));
    int rc = NCB_OK;
    try {
        CK(cudaMemcpy(d, probabilities, bytes, cudaMemcpyHostToDevice));
        rc = ncb_apply_measurement_error_device(d, mats, qubits, nq, num_qubits, nullptr);
        if (rc == NCB_OK) CK(cudaMemcpy(probabilities, d, bytes, cudaMemcpyDeviceToHost));
    } catch (...) { cudaFree(d); throw; }
    cudaFree(d);
    return rc;
    NCB_CATCH
}

// The reference's entry points execute strictly in the caller's instruction order (Simulator.cpp:76-94,
// SimulatorMPI.cpp:40-59): the same mt19937_64 uniform then meets the same state as in the reference.
static ncb_options reference_options() {
    ncb_options o;
    std::memset(&o, 0, sizeof(o));
    o.low_bits = -1; o.fuse = 1; o.device = -1; o.strict_order = 1;
    return o;
}

int ncb_simulate_circuit(const ncb_circuit* circuit, double* statevector, int noisy, uint32_t num_trajectories,
                         int return_state, uint32_t thread_count) {
    (void)thread_count;
    if (!circuit || !statevector) return fail(NCB_E_INVALID, "null argument");
    ncb_program* p = nullptr;
    const ncb_options opt = reference_options();
    int rc = ncb_program_create(circuit, noisy ? NCB_MODE_MCWF : NCB_MODE_PURE, &opt, &p);
    if (rc != NCB_OK) return rc;
    const long long dim = 1ll << circuit->num_qubits;
    if (noisy) {
        std::vector<double> probs(dim, 0.0);
        ncb_run_params rp;
        std::memset(&rp, 0, sizeof(rp));
        rp.traj_begin = 0; rp.traj_count = num_trajectories; rp.rng = NCB_RNG_MT19937_REFERENCE; rp.seed = 42;
        rp.probs = probs.data();
        rc = ncb_mcwf_run(p, &rp);
        if (rc == NCB_OK) {
            const double inv = 1.0 / (double)num_trajectories;            // Simulator.cpp:121, 129
            for (long long j = 0; j < dim; ++j) statevector[2 * j] += probs[j] * inv;
        }
    } else if (return_state) {
        rc = ncb_pure_run(p, statevector, nullptr, nullptr);
    } else {
        std::vector<double> probs(dim);
        rc = ncb_pure_run(p, nullptr, probs.data(), nullptr);
        if (rc == NCB_OK)
            for (long long j = 0; j < dim; ++j) { statevector[2 * j] = probs[j]; statevector[2 * j + 1] = 0.0; }
    }
    ncb_program_destroy(p);
    return rc;
}

int ncb_run_trajectory(const ncb_circuit* circuit, double* statevector, uint32_t seed, uint32_t thread_count) {
    (void)thread_count;
    if (!circuit || !statevector) return fail(NCB_E_INVALID, "null argument");
    ncb_program* p = nullptr;
    const ncb_options opt = reference_options();
    int rc = ncb_program_create(circuit, NCB_MODE_MCWF, &opt, &p);
    if (rc != NCB_OK) return rc;
    const long long dim = 1ll << circuit->num_qubits;
    std::vector<double> probs(dim, 0.0);
    ncb_run_params rp;
    std::memset(&rp, 0, sizeof(rp));
    rp.traj_begin = 0; rp.traj_count = 1; rp.rng = NCB_RNG_MT19937_REFERENCE; rp.seed = seed;
    rp.probs = probs.data();
    rc = ncb_mcwf_run(p, &rp);
    if (rc == NCB_OK)
        for (long long j = 0; j < dim; ++j) { statevector[2 * j] = probs[j]; statevector[2 * j + 1] = 0.0; }
    ncb_program_destroy(p);
    return rc;
}

}  // extern "C"
