"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/ncb200.h
declares, the packer mirrors the reference's lookups and errors, the host-only planner agrees with the oracle's
branch choices, and run calls fail loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from noisycircuits_b200 import _lib, circuits
from noisycircuits_b200.program import PackedCircuit, Program
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ncb200.h")).read()
    declared = set(re.findall(r"\b(ncb_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.lib()
    for sym in declared:
        assert getattr(L, sym) is not None
    assert L.ncb_version() == 100


def test_packer_mirrors_reference_lookups(heron_noise):
    sq, tq, _, _ = heron_noise
    instr = [["sx", [1, 1], 0], ["cz", [1, 0], 0], ["rz", [2], 0.3], ["unitary", [0, 2], np.eye(4)]]
    pk = PackedCircuit(3, instr, sq, tq)
    assert list(pk.opcode[:4]) == [1, 7, 2, 12]
    assert pk.q1[2] == 2                                   # one-element qubit list
    assert list(pk.unitary_qubits[24:26]) == [2, 0]        # reversed, Simulator.cpp:182
    assert pk.channel[3] == -1                             # unitary: no noise (FunctionMapper.hpp:77)
    assert pk.kraus_count[0] == len(sq[1]["sx"]) and pk.kraus_count[1] == len(tq["cz"][(1, 0)])
    with pytest.raises(IndexError):                        # NoiseParser.hpp:137 .at() -> IndexError
        PackedCircuit(3, [["cz", [0, 2], 0]], sq, tq)
    with pytest.raises(ValueError):
        PackedCircuit(3, [["foo", [0, 0], 0]], sq, tq)


def test_compile_errors_are_reported(heron_noise):
    sq, tq, _, _ = heron_noise
    with pytest.raises(_lib.NcbError, match="qubit out of range"):
        Program(PackedCircuit(2, [["sx", [5, 5], 0]], noisy=False), _lib.MODE_PURE)
    with pytest.raises(_lib.NcbError, match="repeated qubit"):
        Program(PackedCircuit(2, [["cz", [1, 1], 0]], noisy=False), _lib.MODE_PURE)


@pytest.mark.parametrize("noise_name", ["heron", "eagle"])
def test_planner_events_agree_with_oracle_choices(heron_noise, eagle_noise, noise_name):
    sq, tq, _, _ = heron_noise if noise_name == "heron" else eagle_noise
    n = 5
    pairs = circuits.coupled_pairs(tq, n)
    instr = (circuits.heron_layered if noise_name == "heron" else circuits.eagle_layered)(n, 3, pairs, seed=4)
    prog = Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF)
    O = oracle.Program(n, instr, sq, tq)
    rng = np.random.default_rng(0)
    base_choice = {}
    n_events = 0
    for t in range(40):
        u = rng.random(len(instr)) if t % 2 else 1 - 0.05 * rng.random(len(instr)) ** 2
        ev = {int(i): (int(a), int(b)) for i, a, b in prog.events(u)}
        _st, choices, _pc = O.trajectory(u, mask_fix=True, diagnostics=True)
        for g in range(len(instr)):
            if g in ev:
                a, b = ev[g]
                assert a <= choices[g] <= b
                n_events += 1
            else:                                   # base branch: one fixed operator per channel
                assert base_choice.setdefault(O.chan[g], choices[g]) == choices[g]
    assert n_events > 0


def test_reference_uniform_stream(heron_noise):
    sq, tq, _, _ = heron_noise
    n = 4
    instr = circuits.heron_ansatz(n, 2, circuits.coupled_pairs(tq, n), seed=1)
    prog = Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF)
    O = oracle.Program(n, instr, sq, tq)
    for seed in (42, 43, 1000):
        assert np.array_equal(prog.reference_uniforms(seed), O.reference_uniforms(seed))


def test_no_cpu_fallback(heron_noise):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    sq, tq, _, _ = heron_noise
    prog = Program(PackedCircuit(2, [["sx", [0, 0], 0]], sq, tq), _lib.MODE_MCWF)
    with pytest.raises(_lib.NcbError) as e:
        prog.run_mcwf(4)
    assert e.value.code == -2          # NCB_E_CUDA


def test_shard_ranges_match_reference_split():
    from noisycircuits_b200.backend import shard_range
    for T in (1, 7, 100, 100000):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                s, c = shard_range(T, r, world)
                base, rem = T // world, T % world                       # QuantumCircuitMPI.py:329-332
                assert (s, c) == (r * base + min(r, rem), base + (1 if r < rem else 0))
                cover += list(range(s, s + c))
            assert cover == list(range(T))


def test_specialised_kernel_sources_compile_for_sm_100a(tmp_path, monkeypatch):
    """ncb_options.specialize: the generator (ncb_spec.cpp) writes one kernel per segment with the pass programs spelled
    out; NVRTC compiles them in process (device headers embedded in the library) here without a GPU.  Checks a lean tiled program end to end up to the cubin cache, and
    that programs the generated kernels do not cover (resident, dense two-qubit operators) are refused cleanly."""
    from noisycircuits_b200 import circuits, noise_io
    from noisycircuits_b200.program import PackedCircuit, Program
    sq, tq, _, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "heron20_noise.npz"))
    monkeypatch.setenv("NCB_SPEC_CACHE", str(tmp_path))
    n = 13
    instr = circuits.heron_layered(n, 3, circuits.coupled_pairs(tq, n), seed=2)
    prog = Program(PackedCircuit(n, instr, sq, tq), tile_bits=9, specialize="parametric")
    src = prog.specialized_source()
    nseg = prog.query(4)
    # per segment: the whole-segment kernel and the kernel of the works that carry jump events
    assert src.count('extern "C" __global__') == 2 * nseg and src.count("_ev(const __grid_constant__ SpecCoef C, DevProgram") == nseg
    assert "spec_diag<" in src and "st_stream(dst" in src
    assert "switch" not in src and "sc." not in src             # no interpreter left in the generated code
    assert "C.c[" in src and "0x1p" not in src                  # coefficients come from the constant bank, not as literals
    # ... so the source (and the cubin cache key) depends on the circuit's structure only: other angles, same kernels
    instr2 = [[g, q, (t * 0.7 + 0.1 if g in ("rx", "rz", "rzz") else t)] for g, q, t in instr]
    prog2 = Program(PackedCircuit(n, instr2, sq, tq), tile_bits=9, specialize="parametric")
    assert prog2.specialized_source() == src
    # specialize=True: coefficients as exact literals (the fastest code; one cubin set per parameter set)
    lit = Program(PackedCircuit(n, instr, sq, tq), tile_bits=9, specialize=True).specialized_source()
    assert "C.c[" not in lit and "0x1." in lit and lit != Program(PackedCircuit(n, instr2, sq, tq), tile_bits=9, specialize=True).specialized_source()
    if not prog.specialize_build():
        pytest.skip("no libnvrtc / nvcc on this machine")
    cubins = [f for d, _, fs in os.walk(tmp_path) for f in fs if f.endswith(".cubin")]
    assert len(cubins) == (nseg + 3) // 4
    assert prog.specialize_build()                              # second call: cache hit
    # a state that fits one CTA: one module with the generated resident kernels (trunk + trajectories), structure-only source
    si = circuits.heron_layered(6, 2, circuits.coupled_pairs(tq, 6), seed=1)
    small = Program(PackedCircuit(6, si, sq, tq), specialize=True)
    ssrc = small.specialized_source()
    assert ssrc.count('extern "C" __global__') == 2 and "ncb_res_trunk" in ssrc and "ncb_res(" in ssrc and "0x1." not in ssrc
    si2 = [[g, q, (t * 0.7 + 0.1 if g in ("rx", "rz", "rzz") else t)] for g, q, t in si]
    assert Program(PackedCircuit(6, si2, sq, tq), specialize="parametric").specialized_source() == ssrc
    assert small.specialize_build()
    assert any(f == "resident.cubin" for d, _, fs in os.walk(tmp_path) for f in fs)


def test_partial_measurement_maps_each_qubit_to_its_bit_of_the_marginal(heron_noise):
    """backend.postprocess on a measured subset that is NOT a permutation of 0..m-1 (the reference strides by the qubit
    id and reads out of bounds there, MeasurementErrorApplicator.cpp:109-112): checked against a dense numpy computation."""
    from noisycircuits_b200.backend import postprocess
    _sq, _tq, meas, _ = heron_noise
    n = 4
    rng = np.random.default_rng(3)
    p = rng.random(1 << n)
    p /= p.sum()
    mdict = {q: meas[q] for q in range(n)}

    def host_meas(probs, noise, qubit_list, num_qubits, num_threads):
        oracle.measurement_error(probs, noise, qubit_list, num_qubits)

    for qubits in ([2, 3], [1], [0, 2], [3, 1], [0, 1, 2, 3]):
        got = postprocess(p.copy(), n, qubits, mdict, 1, host_meas)
        t = p.reshape([2] * n)                                  # axis n-1-q <-> qubit q
        for q in qubits:                                        # dense: read-out matrix of q on q's axis
            t = np.moveaxis(np.tensordot(np.asarray(meas[q], float), t, axes=(1, n - 1 - q)), 0, n - 1 - q)
        kept = sorted(qubits)
        marg = t.sum(axis=tuple(n - 1 - q for q in range(n) if q not in kept))     # axes now: kept qubits, descending
        ref = marg.transpose(list(range(len(kept)))[::-1]).reshape(-1)             # big-endian: lowest kept qubit = MSB
        assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-15, qubits


def test_reference_rng_implies_program_order(heron_noise):
    """rng="reference" (mt19937_64 streams) compiles in strict program order: the reference's branch decisions see the
    state of that order (Simulator.cpp:76-94), a re-ordered program would only match it in distribution."""
    from noisycircuits_b200.solvers import RemoteExecutor
    sq, tq, _, _ = heron_noise
    n = 14
    instr = circuits.heron_layered(n, 3, circuits.coupled_pairs(tq, n), seed=3)
    ex = RemoteExecutor(n, sq, tq, 1, rng="reference")
    assert list(ex._program(instr).order()) == list(range(len(instr)))
    free = RemoteExecutor(n, sq, tq, 1)
    assert list(free._program(instr).order()) != list(range(len(instr)))           # default: tile-friendly order


def test_generated_kernels_compile_with_nvcc_too(tmp_path, monkeypatch):
    """NCB_SPEC_COMPILER=nvcc: the same sources through the nvcc binary (needs the toolkit and csrc/ on disk)."""
    if not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("nvcc not available")
    from noisycircuits_b200 import noise_io
    sq, tq, _, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "heron20_noise.npz"))
    monkeypatch.setenv("NCB_SPEC_CACHE", str(tmp_path))
    monkeypatch.setenv("NCB_SPEC_COMPILER", "nvcc")
    n = 13
    instr = circuits.heron_layered(n, 1, circuits.coupled_pairs(tq, n), seed=2)
    prog = Program(PackedCircuit(n, instr, sq, tq), tile_bits=9, reg_bits=4, specialize=True)      # R = 4 programs too
    assert prog.specialize_build()
    logs = [os.path.join(d, f) for d, _, fs in os.walk(tmp_path) for f in fs if f.endswith(".log")]
    assert logs and "ptxas info" in open(logs[0]).read()


def test_program_cache_is_content_addressed(heron_noise):
    """cache.get_program: the packed bytes (gates, angles, the Kraus operators used) are the key -- a fresh RemoteExecutor
    on the same circuit reuses the compiled program, another angle or another Kraus operator does not."""
    from noisycircuits_b200 import cache
    from noisycircuits_b200.solvers import RemoteExecutor
    sq, tq, _, _ = heron_noise
    n = 6
    instr = circuits.heron_layered(n, 2, circuits.coupled_pairs(tq, n), seed=5)
    cache.clear_programs()
    a = RemoteExecutor(n, sq, tq, 1)._program(instr)
    b = RemoteExecutor(n, sq, tq, 1)._program([list(x) for x in instr])              # new executor, equal content
    assert a is b and cache.stats["program_hits"] >= 1
    other = [list(x) for x in instr]
    other[0] = ["rx", other[0][1], 0.123]
    assert RemoteExecutor(n, sq, tq, 1)._program(other) is not a
    sq2 = {q: {g: [np.array(k) for k in ops] for g, ops in gates.items()} for q, gates in sq.items()}
    g0 = instr[0][0]                                     # a gate the circuit really applies on qubit 0
    sq2[0][g0][0] = sq2[0][g0][0] * 0.999
    assert RemoteExecutor(n, sq2, tq, 1)._program(instr) is not a
    assert RemoteExecutor(n, sq, tq, 1, strict_order=True)._program(instr) is not a   # options are part of the key


def test_program_cache_recycles_engine_objects_for_new_parameter_sets(heron_noise):
    """A training loop calls execute once per parameter set: every call is a cache miss.  Once the cache is full a miss
    recycles the least recently used program of the same kind in place (Program.update) instead of creating an engine object;
    an executor that still holds the recycled object notices by its digest and asks the cache again."""
    from noisycircuits_b200 import cache
    from noisycircuits_b200.solvers import RemoteExecutor
    sq, tq, _, _ = heron_noise
    n = 6
    pairs = circuits.coupled_pairs(tq, n, adjacent_only=True)
    sets = [circuits.qnn_circuit(n, 2, pairs, seed=s) for s in range(cache.PROGRAM_CAPACITY + 3)]
    cache.clear_programs()
    cache.stats["program_recycled"] = 0
    first_holder = RemoteExecutor(n, sq, tq, 1)
    p0 = first_holder._program(sets[0])
    d0 = p0.digest
    ex = RemoteExecutor(n, sq, tq, 1)
    progs = [ex._program(s) for s in sets[1:]]
    assert cache.stats["program_recycled"] == 3 and len({id(p) for p in progs} | {id(p0)}) == cache.PROGRAM_CAPACITY
    assert p0.digest != d0                                   # sets[0]'s object now holds another circuit ...
    again = first_holder._program(sets[0])                   # ... and its first holder does not use it blindly
    assert again.digest == d0 and again.packed.theta.tobytes() == PackedCircuit(n, sets[0], sq, tq).theta.tobytes()
    # the recycled object is the updated program, not a stale one: its order / description belong to the new circuit
    fresh = Program(PackedCircuit(n, sets[-1], sq, tq))
    assert progs[-1].describe() == fresh.describe() and list(progs[-1].order()) == list(fresh.order())


def test_spec_cache_eviction(tmp_path, monkeypatch, heron_noise):
    """The cubin cache keeps at most NCB_SPEC_CACHE_MAX program directories (least recently used go first)."""
    sq, tq, _, _ = heron_noise
    monkeypatch.setenv("NCB_SPEC_CACHE", str(tmp_path))
    monkeypatch.setenv("NCB_SPEC_CACHE_MAX", "2")
    n = 13
    for seed in range(3):
        instr = circuits.heron_layered(n, 1, circuits.coupled_pairs(tq, n), seed=seed)[:6 + seed]
        prog = Program(PackedCircuit(n, instr, sq, tq), tile_bits=9, specialize=True)
        if not prog.specialize_build():
            pytest.skip("no libnvrtc / nvcc on this machine")
    assert len([d for d in os.listdir(tmp_path) if os.path.isdir(os.path.join(tmp_path, d))]) == 2


def test_execute_argument_errors_match_the_reference(heron_noise):
    """backend.execute / execute_sharded raise what QuantumCircuit.execute raises, before anything touches a device
    (QuantumCircuit.py:391-409; the reference's tests/test_QuantumCircuit_error_handling.py pins these)."""
    from noisycircuits_b200.backend import execute, execute_sharded
    sq, tq, meas, _ = heron_noise
    n = 3
    instr = circuits.heron_layered(n, 1, circuits.coupled_pairs(tq, n), seed=1)
    for fn in (execute, execute_sharded):
        with pytest.raises(TypeError, match="Qubits must be a list of integers"):
            fn(n, instr, sq, tq, meas, "012", 10)
        with pytest.raises(TypeError, match="Qubits must be a list of integers"):
            fn(n, instr, sq, tq, meas, [0, 1.5], 10)
        with pytest.raises(ValueError, match="out of range"):
            fn(n, instr, sq, tq, meas, [0, 3], 10)
        with pytest.raises(ValueError, match="No instructions"):
            fn(n, [], sq, tq, meas, [0], 10)
        with pytest.raises(TypeError, match="num_trajectories must be an integer"):
            fn(n, instr, sq, tq, meas, [0], 10.0)
        with pytest.raises(ValueError, match="positive integer"):
            fn(n, instr, sq, tq, meas, [0], 0)


def test_sweep_pipeline_never_touches_an_engine_in_flight(heron_noise, monkeypatch):
    """RemoteExecutor.run_sweep / run_many (host logic, engine objects replaced by stand-ins): circuit i is launched before
    circuit i - 1 is collected, preparations run on host threads into engine objects that are neither in flight nor being
    prepared, results come back in order, and an exception leaves no run in flight."""
    import threading
    import time
    from noisycircuits_b200 import solvers
    sq, tq, _, _ = heron_noise
    n = 5
    lock = threading.Lock()
    log = {"max_in_flight": 0, "in_flight": 0, "engines": 0, "updates": 0}

    class FakeProgram:
        def __init__(self, packed, mode, **opts):
            with lock:
                log["engines"] += 1
            self.busy, self.preparing, self.packed = False, False, packed
            time.sleep(0.002)

        def update(self, packed):
            assert not self.busy, "an engine was updated while its run was in flight"
            assert not self.preparing, "two preparations in one engine"
            self.preparing = True
            time.sleep(0.002)
            self.packed = packed
            self.preparing = False
            with lock:
                log["updates"] += 1

        def launch_mcwf(self, traj_count, **kw):
            assert not self.busy and not self.preparing
            self.busy = True
            with lock:
                log["in_flight"] += 1
                log["max_in_flight"] = max(log["max_in_flight"], log["in_flight"])
            if float(self.packed.theta[0]) == 666.0:
                self.busy = False
                with lock:
                    log["in_flight"] -= 1
                raise solvers._lib.NcbError(1, "boom")
            return (self, float(self.packed.theta[0]))

        def collect_mcwf(self, handle):
            assert self.busy
            self.busy = False
            with lock:
                log["in_flight"] -= 1
            return np.full(1 << n, handle[1] * traj), {"events": 0}

    traj = 10
    monkeypatch.setattr(solvers, "Program", FakeProgram)
    skeleton = [["rx", [q, q], 0.0] for q in range(n)]
    thetas = np.zeros((23, n))
    thetas[:, 0] = np.arange(23) + 1.0
    ex = solvers.RemoteExecutor(n, sq, tq, 1)
    out = ex.run_sweep(traj, skeleton, thetas)
    assert [float(o[0]) for o in out] == [float(k + 1) for k in range(23)]                 # in order, divided by the trajectory count
    assert log["max_in_flight"] == 2 and log["in_flight"] == 0                             # one running, one queued behind it
    assert log["engines"] == ex.PREPARE_THREADS + 2 and log["updates"] == 23 - log["engines"]
    lists = [[["rx", [q, q], float(k + 1) if q == 0 else 0.0] for q in range(n)] for k in range(9)]
    assert [float(o[0]) for o in ex.run_many(traj, lists)] == [float(k + 1) for k in range(9)]
    assert log["engines"] == ex.PREPARE_THREADS + 2                                          # the engine objects are kept between sweeps
    thetas[7, 0] = 666.0
    with pytest.raises(solvers._lib.NcbError):
        ex.run_sweep(traj, skeleton, thetas)
    assert log["in_flight"] == 0                                                            # nothing left in flight
    assert len(ex.run_sweep(traj, skeleton, thetas[:5])) == 5                               # and the executor still works (fresh engines)


def test_next_parameter_set_reuses_the_packed_gate_skeleton(heron_noise):
    """RemoteExecutor._program on the last circuit's gate skeleton with other parameters replaces the parameter array only;
    the packed circuit (and with it the content key of the program cache) is what packing from scratch gives."""
    from noisycircuits_b200 import cache
    from noisycircuits_b200.solvers import RemoteExecutor, _fingerprint
    sq, tq, _, _ = heron_noise
    n = 6
    pairs = circuits.coupled_pairs(tq, n, adjacent_only=True)
    a, b = circuits.qnn_circuit(n, 2, pairs, seed=1), circuits.qnn_circuit(n, 2, pairs, seed=2)
    ex = RemoteExecutor(n, sq, tq, 1)
    assert ex._pack(_fingerprint(a)) is None                          # nothing known yet
    ex._program(a)
    reused = ex._pack(_fingerprint(b))
    fresh = PackedCircuit(n, b, sq, tq)
    assert reused is not None
    for name in ("opcode", "q0", "q1", "theta", "channel", "unitary_offset", "unitary_nq", "unitary_qubits", "unitary_pool",
                 "channel_offset", "channel_count", "channel_dim", "kraus_pool"):
        assert getattr(reused, name).tobytes() == getattr(fresh, name).tobytes(), name
    assert cache.packed_digest(reused, 0, {"specialize": False}) == cache.packed_digest(fresh, 0, {"specialize": False})
    assert cache.packed_digest(reused, 0, {"specialize": False}) != cache.packed_digest(PackedCircuit(n, a, sq, tq), 0, {"specialize": False})
    # another skeleton (one gate more), a unitary, or other noise dictionaries: packed from scratch
    assert ex._pack(_fingerprint(b + [["x", [0, 0], 0]])) is None
    assert ex._pack(_fingerprint(b[:-1] + [["unitary", [0], np.eye(2)]])) is None
    ex2 = RemoteExecutor(n, dict(sq), tq, 1)
    assert ex2._pack(_fingerprint(b)) is None
    p1 = ex._program(b)
    assert p1.packed.theta.tobytes() == fresh.theta.tobytes() and p1.describe() == Program(fresh).describe()
