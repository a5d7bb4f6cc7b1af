"""
Content-addressed caches: what the reference redoes on every construction / call is done once.

  * compiled programs, in process: the reference re-parses the instruction list and every Kraus dictionary on every
    ``simulate_circuit`` call (Simulator.cpp:162-196) and builds a new executor per ``execute`` (QuantumCircuit.py:411-416).
    Here a call packs the circuit (milliseconds) and looks the PACKED BYTES up -- gate list, angles, the Kraus operators
    actually used -- so a repeated ``execute`` of the same circuit, also through a fresh ``RemoteExecutor``, reuses the
    compiled program, its uploaded tables and its loaded generated kernels and is launch bound.
  * generated kernels, on disk: the cubin cache of ncb_engine.cu (keyed by a hash of the generated source).
  * noise operators, on disk: ``QuantumCircuit.__init__`` runs ``BuildModel.build_qubit_gate_model`` (joblib, seconds) on
    every construction (QuantumCircuit.py:149-157).  ``cached_qubit_gate_model`` keys its result by the bytes of the noise
    dictionary and the build arguments and stores it under the cache directory; ``install(cache_noise_tables=True)`` puts it
    in front of the reference's builder.

Cache directory: $NCB_CACHE_DIR, else ~/.cache/noisycircuits_b200 (a temporary directory when that is not writable).
"""
from __future__ import annotations

import collections
import hashlib
import os
import pickle
import tempfile

from . import _lib

PROGRAM_CAPACITY = int(os.environ.get("NCB_PROGRAM_CACHE", "4"))      # programs hold their state batch on the device
_programs: "collections.OrderedDict[str, object]" = collections.OrderedDict()
stats = {"program_hits": 0, "program_misses": 0, "noise_hits": 0, "noise_misses": 0}


def packed_digest(packed, mode: int, options: dict) -> str:
    """blake2b of everything a compiled program depends on: the packed circuit (ncb_circuit arrays), mode, options."""
    # everything but the gate parameters is hashed once per gate skeleton (PackedCircuit.with_thetas hands the digest on): the
    # Kraus pool is a few hundred KB, the parameters of a training loop's next circuit a few KB
    static = getattr(packed, "_static_digest", None)
    if static is None:
        hs = hashlib.blake2b(digest_size=20)
        hs.update(repr((int(packed.num_qubits), int(packed.num_instructions), int(packed.num_channels))).encode())
        for a in (packed.opcode, packed.q0, packed.q1, packed.channel, packed.unitary_offset, packed.unitary_nq,
                  packed.unitary_qubits, packed.unitary_pool, packed.channel_offset, packed.channel_count, packed.channel_dim, packed.kraus_pool):
            hs.update(a.tobytes())
        static = packed._static_digest = hs.digest()
    h = hashlib.blake2b(digest_size=20)
    h.update(static)
    h.update(repr((int(mode), sorted((k, repr(v)) for k, v in options.items()))).encode())
    h.update(packed.theta.tobytes())
    return h.hexdigest()


def get_program(packed, mode: int = _lib.MODE_MCWF, **options):
    """The compiled ``Program`` of this packed circuit (compiled on the first request, LRU of PROGRAM_CAPACITY).

    ``prog.digest`` names the circuit the object holds.  When the cache is full, a miss RECYCLES the least recently used
    program of the same kind (mode, options, qubits) with ``Program.update`` instead of building a new engine object: a
    training loop that calls ``execute`` once per parameter set then pays a re-compilation (under a millisecond) per call, not
    device allocations and a module load.  Holders of a program from an earlier call check ``digest`` before reusing it
    (``RemoteExecutor._program`` does)."""
    from .program import Program
    key = packed_digest(packed, mode, options)
    prog = _programs.get(key)
    if prog is not None:
        _programs.move_to_end(key)
        stats["program_hits"] += 1
        return prog
    stats["program_misses"] += 1
    kind = (int(mode), repr(sorted((k, repr(v)) for k, v in options.items())), int(packed.num_qubits))
    prog = None
    if len(_programs) >= max(PROGRAM_CAPACITY, 1):
        for old_key, cand in _programs.items():              # oldest first
            if getattr(cand, "kind", None) == kind:
                try:
                    cand.update(packed)
                except _lib.NcbError:                         # another geometry (tile / register bits): a new object after all
                    break
                del _programs[old_key]
                prog = cand
                stats["program_recycled"] = stats.get("program_recycled", 0) + 1
                break
    if prog is None:
        prog = Program(packed, mode, **options)
        prog.kind = kind
    prog.digest = key
    _programs[key] = prog
    while len(_programs) > max(PROGRAM_CAPACITY, 1):
        _programs.popitem(last=False)          # (dropped, not closed: an executor may still hold it)
    return prog


def clear_programs() -> None:
    _programs.clear()


def cache_dir() -> str:
    d = os.environ.get("NCB_CACHE_DIR") or os.path.join(os.path.expanduser("~"), ".cache", "noisycircuits_b200")
    try:
        os.makedirs(d, exist_ok=True)
        if os.access(d, os.W_OK):
            return d
    except OSError:
        pass
    d = os.path.join(tempfile.gettempdir(), "noisycircuits_b200")
    os.makedirs(d, exist_ok=True)
    return d


def noise_model_digest(noise_model: dict, num_qubits: int, threshold: float, basis_gates) -> str:
    h = hashlib.blake2b(digest_size=20)
    h.update(pickle.dumps((int(num_qubits), float(threshold), [list(g) for g in basis_gates]), protocol=4))
    h.update(pickle.dumps(noise_model, protocol=4))
    return h.hexdigest()


def cached_qubit_gate_model(modeller, build):
    """``build(modeller)`` (= BuildModel.build_qubit_gate_model, BuildQubitGateModel.py:852-904) through the on-disk cache.
    Returns the builder's tuple (single, two, measurement, connectivity); the attribute the constructor reads afterwards
    (``qubit_coupling_map``, QuantumCircuit.py:175) is restored on the modeller too."""
    key = noise_model_digest(modeller.noise_model, modeller.num_qubits, modeller.threshold, modeller.basis_gates)
    path = os.path.join(cache_dir(), "noise-" + key + ".pkl")
    if os.path.exists(path):
        try:
            with open(path, "rb") as f:
                blob = pickle.load(f)
            modeller.qubit_coupling_map = blob["qubit_coupling_map"]
            stats["noise_hits"] += 1
            return blob["result"]
        except Exception:
            pass                                # unreadable entry: rebuild
    stats["noise_misses"] += 1
    result = build(modeller)
    tmp = path + ".%d.tmp" % os.getpid()
    try:
        with open(tmp, "wb") as f:
            pickle.dump({"result": result, "qubit_coupling_map": getattr(modeller, "qubit_coupling_map", None)}, f, protocol=4)
        os.replace(tmp, path)
    except OSError:
        pass
    return result
