/*
 * ncb200.h -- C ABI of the B200 noisy-circuit execution engine (libncb200.so).
 *
 * Drop-in boundary.  The reference package NoisyCircuits binds its native hot path as three pybind11 modules
 * (paths relative to /root/reference/src/NoisyCircuits/utils/custom/src/):
 *
 *   simulator.simulate_circuit(instructions, statevector, single_qubit_noise_instructions,
 *                              two_qubit_noise_instructions, num_qubits, noisy, num_trajectories,
 *                              return_state, thread_count)                         Simulator.cpp:161, 209-212
 *   simulator_mpi.run_trajectory(instructions, statevector, single_qubit_instructions,
 *                                two_qubit_instructions, num_qubits, seed, thread_count)
 *                                                                                  SimulatorMPI.cpp:83, 124-127
 *   measurement_error_applicator.apply_measurement_error(probabilities_array, measurement_noise, qubit_list,
 *                                                        num_qubits, num_threads)
 *                                                                     MeasurementErrorApplicator.cpp:101, 119-122
 *
 * Python lists / dicts cannot cross a C ABI, so the instruction list and the Kraus dictionaries are passed in the
 * packed form `ncb_circuit` (what Simulator.cpp:162-196 + NoiseParser.hpp:52-106 unpack on every call);
 * ncb_simulate_circuit / ncb_run_trajectory / ncb_apply_measurement_error are the three entry points above with
 * that packing, same argument meaning, same in/out buffer conventions.  The ncb_program_* / ncb_*_run functions are
 * the compile-once / run-many form the Python solver classes (noisycircuits_b200.solvers) use.
 *
 * All pointers are host pointers unless the name says `device`.  Complex data is interleaved (re, im) float64.
 * Every function returns 0 on success and a negative NCB_E_* code on failure; ncb_last_error() gives the message
 * (thread-local).  There is no CPU fallback: run functions fail with NCB_E_CUDA when no CUDA device is usable.
 */
#ifndef NCB200_H
#define NCB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NCB_VERSION 100

enum { NCB_OK = 0, NCB_E_INVALID = -1, NCB_E_CUDA = -2, NCB_E_UNSUPPORTED = -3, NCB_E_NOMEM = -4 };

/* gate opcodes: the gate names of FunctionMapper.hpp:27-39 */
enum {
    NCB_X = 0, NCB_SX = 1, NCB_RZ = 2, NCB_RX = 3, NCB_RY = 4, NCB_H = 5, NCB_P = 6,
    NCB_CZ = 7, NCB_RZZ = 8, NCB_ECR = 9, NCB_CX = 10, NCB_SWAP = 11, NCB_UNITARY = 12
};

/* Packed (instruction list, noise dictionaries).
 *   instruction g: opcode[g]; single-qubit gates: q0[g] (q1 ignored); two-qubit gates: q0[g], q1[g] in the order
 *   the instruction lists them (matrix index = bit(q0) + 2*bit(q1), QuantumGates.hpp:209-219); theta[g].
 *   NCB_UNITARY: unitary_nq[g] target qubits unitary_qubits[8*g + b] for matrix bit b (i.e. the instruction's qubit
 *   list REVERSED, Simulator.cpp:182), row-major 2^k x 2^k matrix at unitary_pool + 2*unitary_offset[g]; no noise.
 *   channel[g]: index of the Kraus list applied after the gate, or -1 (no noise step).  The Python side resolves
 *   single_qubit_noise[q][gate] / two_qubit_noise[gate][(q0,q1)] to channel ids (NoiseParser.hpp:130-140).
 *   channel c: channel_count[c] operators of size channel_dim[c] (2 or 4), row-major, at
 *   kraus_pool + 2*channel_offset[c]. */
typedef struct {
    int32_t num_qubits;
    int32_t num_instructions;
    const int32_t* opcode;
    const int32_t* q0;
    const int32_t* q1;
    const double* theta;
    const int32_t* channel;
    const int64_t* unitary_offset;
    const int32_t* unitary_nq;
    const int32_t* unitary_qubits;
    const double* unitary_pool;
    int32_t num_channels;
    const int64_t* channel_offset;
    const int32_t* channel_count;
    const int32_t* channel_dim;
    const double* kraus_pool;
} ncb_circuit;

enum { NCB_MODE_MCWF = 0, NCB_MODE_DENSITY = 1, NCB_MODE_PURE = 2 };

typedef struct {
    int32_t tile_bits;   /* amplitudes per shared-memory tile = 2^tile_bits (0: default 11) */
    int32_t reg_bits;    /* amplitudes per thread = 2^reg_bits (0: default 3; 4 is needed for 3-4 qubit unitaries and density-matrix runs) */
    int32_t low_bits;    /* lowest state bits kept in every tile (-1: default 2) */
    int32_t fuse;        /* 1 (default): fuse runs of instructions into one sweep; 0: one sweep per instruction */
    int32_t device;      /* CUDA device ordinal (-1: current device) */
    int32_t strict_order;/* 0 (default): commuting instructions (disjoint qubits) may be executed in a tile-friendly
                            order -- the same circuit, fewer sweeps; jump events then have the reference's meaning on
                            that order (ncb_program_order).  1: execute strictly in the caller's order. */
    int32_t independent_trajectories;
                         /* 0 (default): the batched MCWF run computes the evolution every trajectory shares before its
                            first jump event once per batch (results per trajectory are bit-identical; fewer sweeps).
                            1: every trajectory is swept on its own from |0..0>. */
    int32_t specialize;  /* 1 or 2: the sweeps run in kernels GENERATED for this program (per segment one kernel for the works that
                            run the whole segment and one for the works that carry jump events: the pass programs written out,
                            compiled with nvcc into a cubin cache next to the library, loaded through the driver API) instead
                            of the interpreting kernels -- same arithmetic, a third of the instructions.  1: gate coefficients
                            are literals in the generated code (fastest; the cubins belong to this circuit and these angles);
                            2: coefficients are a kernel parameter, so the generated code -- and the cache key -- depends on
                            the circuit's STRUCTURE only: a parameter sweep over one gate skeleton compiles once.  Costs a few
                            seconds at the first run of a new program; when nvcc or libcuda are not usable the interpreter
                            runs and ncb_run_params.out_spec_launches stays 0.  0 (default): off. */
} ncb_options;

typedef struct ncb_program ncb_program;

const char* ncb_last_error(void);
int ncb_version(void);

/* Compilation is host-only (no CUDA call is made until a run function is called). */
int ncb_program_create(const ncb_circuit* circuit, int mode, const ncb_options* options, ncb_program** out);
void ncb_program_destroy(ncb_program* program);
/* Replaces the program's circuit in place (same mode and options): the new circuit is compiled on the host and takes over
 * the engine object -- device buffers, streams and, when the generated code does not change (ncb_options.specialize = 2: same
 * gate skeleton, other angles), the loaded kernels are kept, so a parameter sweep (QNN training loop) pays a host compile
 * and a table upload per circuit instead of a new engine.  Fails with NCB_E_INVALID when the new circuit compiles to another
 * geometry (number of qubits, tile or register bits). */
int ncb_program_update(ncb_program* program, const ncb_circuit* circuit);
/* Human-readable plan (segments, tiles, passes) into buf; returns the number of bytes needed. */
int64_t ncb_program_describe(const ncb_program* program, char* buf, int64_t len);

enum {
    NCB_Q_NUM_QUBITS = 0, NCB_Q_STATE_BITS = 1, NCB_Q_TILE_BITS = 2, NCB_Q_REG_BITS = 3, NCB_Q_NUM_SEGMENTS = 4,
    NCB_Q_NUM_PASSES = 5, NCB_Q_NUM_OPS = 6, NCB_Q_NUM_DRAWS = 7, NCB_Q_NORM_PRESERVING = 8, NCB_Q_NUM_INSTRUCTIONS = 9,
    NCB_Q_RESIDENT = 10
};
int64_t ncb_program_query(const ncb_program* program, int what);

/* Generates and compiles the specialised kernels of the program into the cubin cache (host only; what the first run
 * with ncb_options.specialize would do).  Returns NCB_E_INVALID when the program has no segment that qualifies or nvcc
 * failed. */
int ncb_program_spec_build(ncb_program* program);
/* CUDA source of the specialised kernels of segments [seg_begin, seg_end); returns the bytes needed. */
int64_t ncb_program_spec_source(const ncb_program* program, int32_t seg_begin, int32_t seg_end, char* buf, int64_t len);

/* Execution order: out[pos] = index (in the caller's instruction list) of the instruction executed at position pos. */
int ncb_program_order(const ncb_program* program, int32_t* out);
/* Jump events of one trajectory (host-only; used by tests): for explicit uniforms u[num_instructions] (indexed like the
 * caller's instruction list) writes up to cap records (instr, kmin, kmax) into out[3*i..], in execution order, and
 * returns the number of events. */
int64_t ncb_program_events(const ncb_program* program, const double* uniforms, int32_t* out, int64_t cap);
/* The reference's uniform stream for std::mt19937_64(seed): u[g] for every instruction that consumes a draw
 * (Kraus list with >= 2 operators), 0.5 elsewhere. */
int ncb_program_reference_uniforms(const ncb_program* program, uint64_t seed, double* uniforms);

enum { NCB_RNG_PHILOX = 0, NCB_RNG_MT19937_REFERENCE = 1, NCB_RNG_EXPLICIT = 2 };

typedef struct {
    int64_t traj_begin;        /* global id of the first trajectory (keys the RNG: results do not depend on sharding) */
    int64_t traj_count;
    int32_t rng;               /* NCB_RNG_* */
    int32_t probs_on_device;   /* probs is a device pointer */
    uint64_t seed;             /* PHILOX: key; MT19937_REFERENCE: trajectory t uses mt19937_64(seed + t) (Simulator.cpp:124) */
    const double* uniforms;    /* EXPLICIT: [traj_count][num_instructions] */
    double* probs;             /* [2^n]: sum over the trajectories of |psi|^2 (NOT divided by the count) */
    int32_t accumulate;        /* 1: add to probs, 0: overwrite */
    int32_t batch;             /* trajectories resident at once (0: automatic) */
    double* states;            /* optional [traj_count][2^n] complex: final normalised state of each trajectory */
    void* stream;              /* cudaStream_t (NULL: default stream); the engine forks an internal second stream from it
                                  and joins it back, so the caller only ever orders against this one */
    /* statistics, filled on return */
    int64_t out_sweeps;        /* state sweeps issued (one read + one write of one trajectory's state each), the shared
                                  no-jump prefix counted once per batch */
    int64_t out_events;        /* jump events (branch != base branch) */
    int64_t out_hard_events;   /* events whose branch needed the reduced density matrix */
    int64_t out_launches;      /* kernel launches */
    double out_kernel_ms;      /* device time from the first to the last kernel of the run (CUDA events on `stream`) */
    int64_t out_tile_launches; /* launches of the sweep (direct / tile / resident) kernels among out_launches */
    int64_t out_h2d_bytes;     /* bytes copied host -> device during the run (work lists, uniform tables) */
    int64_t out_d2h_bytes;     /* bytes copied device -> host during the run (probabilities, states, counters) */
    int64_t out_spec_launches; /* sweep launches that ran in a kernel GENERATED for the program (ncb_options.specialize); 0 with
                                  specialize = 1 means the generated kernels could not be built or loaded and the interpreting
                                  kernels did the work */
} ncb_run_params;

int ncb_mcwf_run(ncb_program* program, ncb_run_params* params);
/* ncb_mcwf_run in two phases, for callers that keep several programs in flight (a parameter sweep: the reference's training
 * loops call QuantumCircuit.execute once per parameter set, examples/quantum_neural_networks.ipynb cells 20-27):
 * ncb_mcwf_launch enqueues the run -- with stream = 0 on a stream of the program's own, so that launches of different programs
 * overlap on the device -- and returns without waiting where the path allows it (programs whose state fits one CTA running
 * generated kernels with Philox draws: their planning happens on the device; every other program has finished its kernels when
 * the call returns); ncb_mcwf_collect, with the SAME params object, waits, fills probs (host buffers are written only here) and
 * the out_* statistics.  One run in flight per program; the buffers of params stay alive in between. */
int ncb_mcwf_launch(ncb_program* program, ncb_run_params* params);
int ncb_mcwf_collect(ncb_program* program, ncb_run_params* params);
/* Noise-free run from |0..0>: state (complex, 2^n) and/or probs (2^n) may be NULL. */
int ncb_pure_run(ncb_program* program, double* state, double* probs, void* stream);
/* Density-matrix run: probs[r] = rho[r][r], r over all 2^n basis states (little-endian). */
int ncb_density_run(ncb_program* program, double* probs, void* stream);

/* In-place read-out error on a probability vector of 2^num_qubits entries: for k in [0, nq): 2x2 row-major
 * matrix mats[4k..] on pairs at stride 1 << qubits[k] (MeasurementErrorApplicator.cpp:65-81, 101-114). */
int ncb_apply_measurement_error(double* probabilities, const double* mats, const int32_t* qubits, int32_t nq,
                                int32_t num_qubits, uint32_t num_threads);
int ncb_apply_measurement_error_device(double* probabilities_device, const double* mats, const int32_t* qubits,
                                       int32_t nq, int32_t num_qubits, void* stream);

/* The tail of QuantumCircuit.execute (QuantumCircuit.py:436-445) without leaving the device:
 *   probabilities_device float64[2^num_qubits], little-endian over all qubits (e.g. the accumulator ncb_mcwf_run filled
 *   with probs_on_device, or the all-reduced sum of the ranks), is multiplied by `scale`, marginalised over `qubits`
 *   (HelperFunctions.py:9-32; kept qubits in ascending order), the read-out error mats[4k..] of measured qubit qubits[k]
 *   is applied on that qubit's bit of the marginal (its rank among the kept qubits -- for "all qubits in order" exactly
 *   what ncb_apply_measurement_error does; the reference strides by the qubit id and is out of bounds for other subsets,
 *   MeasurementErrorApplicator.cpp:109-112), the index order is reversed to big-endian, and the float64[2^nq] result is
 *   copied to the host buffer `out`.  mats may be NULL (no read-out error). */
int ncb_execute_tail_device(const double* probabilities_device, int32_t num_qubits, const int32_t* qubits, int32_t nq,
                            const double* mats, double scale, double* out, void* stream);
/* Device scratch for callers without a CUDA allocator of their own (the Python mirror's execute()). */
int ncb_device_alloc(int64_t bytes, void** out_device_pointer);
int ncb_device_free(void* device_pointer);

/* The reference's FFI, packed (see the header comment).  statevector: complex128[2^n], caller-owned.
 *   noisy != 0: statevector[i] += mean_t |psi_t[i]|^2 (real part), trajectories seeded 42 + t (Simulator.cpp:118-133);
 *   noisy == 0: statevector := final state (return_state != 0) or its probabilities in the real part. */
int ncb_simulate_circuit(const ncb_circuit* circuit, double* statevector, int noisy, uint32_t num_trajectories,
                         int return_state, uint32_t thread_count);
/* One trajectory seeded with `seed`: statevector := |psi|^2 in the real part (SimulatorMPI.cpp:40-59). */
int ncb_run_trajectory(const ncb_circuit* circuit, double* statevector, uint32_t seed, uint32_t thread_count);

#ifdef __cplusplus
}
#endif
#endif
