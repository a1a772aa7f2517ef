"""Commitment traces of whole proofs (BASELINE configs 2, 4, 5).

The Rust prover cannot be built in this image, so a proof is *replayed* as the sequence of
polynomial-batch commitments and FRI-layer trees its ``prove()`` calls make, on synthetic data of the
right shapes.  Every ``prove()`` of degree ``n`` under ``standard_recursion_config``
(mp2-common/src/lib.rs:45-47: 135 wires, 2 challenges, quotient degree factor 8, rate_bits 3,
cap_height 4, ConstantArityBits(4, 5)) commits:

    from_values(135 x n)   wires
    from_values( 20 x n)   Z / partial products  (2 * (1 + 9))
    from_coeffs( 16 x n)   quotient chunks       (2 * 8)
    MerkleTree::new per FRI reduction layer: (8n / 16) leaves of 16*D = 32 elements, then /16 per layer

A proof of one mp2 circuit is a base ``prove()`` followed by the wrap chain down to 2^12 rows
(recursion-framework/src/universal_verifier_gadget/wrap_circuit.rs:64-115, RECURSION_THRESHOLD = 12 at
recursion-framework/src/universal_verifier_gadget/mod.rs:34).  The degrees below are ASSUMED (base 2^14,
wraps 2^13 and 2^12) until a Rust toolchain can report the real ones; a trace's proofs/s is an UPPER
BOUND on prover throughput (witness generation, quotient evaluation and openings are outside the path).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

RATE_BITS, CAP_HEIGHT, NUM_WIRES, ZS_PP_COLS, QUOTIENT_COLS, FRI_LEAF_LEN = 3, 4, 135, 20, 16, 32
ARITY_BITS, FINAL_POLY_BITS = 4, 5

TRACE_INCLUDES = "per prove(): from_values(135 cols), from_values(20), from_coeffs(16), one Merkle tree per FRI layer"

LEAF_PROOF_DEGREES = (14, 13, 12)          # assumed: base circuit + two wrap steps (BASELINE config 2)
AGGREGATION_DEGREES = (13, 12)             # assumed: 2-proof branch circuit + one wrap (BASELINE config 4)


@dataclass(frozen=True)
class Op:
    kind: str       # "from_values" | "from_coeffs" | "merkle"
    ncols: int      # columns, or leaf length for "merkle"
    n_log: int      # log2 rows, or log2 leaves for "merkle"

    @property
    def lde_elems(self) -> int:
        return self.ncols << (self.n_log + RATE_BITS) if self.kind != "merkle" else 0

    @property
    def perms(self) -> int:
        if self.kind == "merkle":
            nl = 1 << self.n_log
            return nl * ((self.ncols + 7) // 8) + nl - (1 << min(CAP_HEIGHT, self.n_log))
        N = 1 << (self.n_log + RATE_BITS)
        return N * ((self.ncols + 7) // 8) + N - (1 << CAP_HEIGHT)


def fri_reduction_arity_bits(degree_bits: int) -> List[int]:
    """plonky2 FriReductionStrategy::ConstantArityBits(4, 5).reduction_arity_bits(...)."""
    out, db = [], degree_bits
    while db > FINAL_POLY_BITS and db + RATE_BITS - CAP_HEIGHT > ARITY_BITS:
        out.append(ARITY_BITS)
        db -= ARITY_BITS
    return out


def prove_ops(degree_bits: int) -> List[Op]:
    ops = [Op("from_values", NUM_WIRES, degree_bits), Op("from_values", ZS_PP_COLS, degree_bits),
           Op("from_coeffs", QUOTIENT_COLS, degree_bits)]
    lde_bits = degree_bits + RATE_BITS
    for ab in fri_reduction_arity_bits(degree_bits):
        lde_bits -= ab
        ops.append(Op("merkle", FRI_LEAF_LEN, lde_bits))
    return ops


def proof_ops(degrees=LEAF_PROOF_DEGREES) -> List[Op]:
    return [op for d in degrees for op in prove_ops(d)]


class TraceRunner:
    """Replays proof traces on the current CUDA device, ``nstreams`` proofs in flight."""

    def __init__(self, degrees=LEAF_PROOF_DEGREES, hash_kind: int = 1, nstreams: int = 8):
        import torch

        from . import device as D

        self.torch, self.D = torch, D
        D.bind_current_device()
        self.ops = proof_ops(degrees)
        self.hash_kind = hash_kind
        self.streams = [torch.cuda.Stream() for _ in range(nstreams)]
        self.slots = [self._alloc() for _ in range(nstreams)]

    def _alloc(self):
        torch, D = self.torch, self.D
        bufs = []
        for op in self.ops:
            if op.kind == "merkle":
                nl = 1 << op.n_log
                cap_h = min(CAP_HEIGHT, op.n_log)
                leaves = torch.randint(0, 1 << 62, (nl, op.ncols), dtype=torch.int64, device="cuda")
                dig = torch.empty((max(2 * (nl - (1 << cap_h)), 1), 4), dtype=torch.int64, device="cuda")
                cap = torch.empty((1 << cap_h, 4), dtype=torch.int64, device="cuda")
                bufs.append((leaves, dig, cap, cap_h))
            else:
                cols = torch.randint(0, 1 << 62, (op.ncols, 1 << op.n_log), dtype=torch.int64, device="cuda")
                bufs.append((cols, D.CommitBuffers(op.ncols, op.n_log, RATE_BITS, CAP_HEIGHT, True)))
        return bufs

    def _one(self, slot):
        D = self.D
        for op, b in zip(self.ops, slot):
            if op.kind == "merkle":
                D.merkle_rowmajor(b[0], b[3], self.hash_kind, b[1], b[2])
            else:
                D.commit_resident(b[0], b[1], self.hash_kind, op.kind == "from_coeffs")

    def run(self, nproofs: int) -> None:
        """Launches ``nproofs`` traces round-robin over the streams (asynchronous)."""
        torch = self.torch
        for i in range(nproofs):
            with torch.cuda.stream(self.streams[i % len(self.streams)]):
                self._one(self.slots[i % len(self.slots)])

    def sync(self) -> None:
        for s in self.streams:
            s.synchronize()

    @property
    def perms_per_proof(self) -> int:
        return sum(op.perms for op in self.ops)

    @property
    def lde_elems_per_proof(self) -> int:
        return sum(op.lde_elems for op in self.ops)


# ------------------------------------------------------------------------------------------------------------------
# Whole device-side prover replay: every step of plonky2's prove() that this library implements, through the
# host-buffer API (the call a patched plonky2 makes: mp2gpu_prove, one native call per prove(); `native=False` replays
# the same sequence through the Python mirror and can log per-stage times), one prover per host thread (per-thread
# streams; ctypes drops the GIL inside every call).  Still a replay on synthetic data: witness generation is host
# work in the reference and is not counted.
# ------------------------------------------------------------------------------------------------------------------
PROVER_INCLUDES = ("per prove() = one mp2gpu_prove call: from_values(135 wires) | challenger | Z / partial products on the "
                   "device + from_values(20 columns) | "
                   "compute_quotient_polys on the device over the 14-kind recursion gate set (trace_circuit_desc) + from_coeffs(16 chunks) | "
                   "OpeningSet evaluations at zeta, g*zeta | "
                   "prove_openings: alpha-batched quotient, FRI commit phase, PoW grind (16 bits), 28 query rounds with "
                   "Merkle paths | bincode(ProofWithPublicInputs); pinned host wire columns in; proof bytes out (rows, "
                   "coefficients and digests stay in HBM)")
NUM_ROUTED_WIRES = 80


def trace_circuit_desc(degree_bits: int, recursion_gate_set: bool = True):
    """A circuit descriptor of the standard_recursion_config shape (135 wires, 80 routed, 2 challenges, quotient
    degree factor 8).  With ``recursion_gate_set`` (default) it carries the 14 gate kinds a recursive verifier circuit
    is built from, at that config's sizes (ArithmeticGate 20 ops, ArithmeticExtension 10, MulExtension 13, BaseSum<2> 63
    limbs, Reducing 43, ReducingExtension 32, RandomAccess 4 bits x 4 copies + 2 constants, Exponentiation 66 bits,
    CosetInterpolation 16 points degree 6, Poseidon, PoseidonMds, Constant, PublicInput, Noop) in five selector groups
    whose filtered degrees stay <= 9; otherwise the five-gate subset round 2 started with.  Every gate's constraints are
    evaluated at every point of the quotient coset, so the set decides the quotient kernel's cost."""
    from .quotient import CircuitDesc, GateDesc

    if not recursion_gate_set:
        return CircuitDesc(degree_bits, NUM_WIRES, NUM_ROUTED_WIRES, 4,
                           [GateDesc("arithmetic", NUM_ROUTED_WIRES // 4), GateDesc("constant", 2), GateDesc("noop"),
                            GateDesc("public_input"), GateDesc("poseidon")], [0, 0, 0, 0, 1], [(0, 4), (4, 5)], 3, 2)
    gates = [GateDesc("noop"), GateDesc("constant", 2), GateDesc("public_input"), GateDesc("poseidon_mds"),
             GateDesc("base_sum", 63, 2),
             GateDesc("arithmetic", 20), GateDesc("arithmetic_extension", 10), GateDesc("mul_extension", 13),
             GateDesc("reducing", 43), GateDesc("reducing_extension", 32),
             GateDesc("random_access", 4, 4 | (2 << 8)), GateDesc("exponentiation", 66),
             GateDesc("coset_interpolation", 4, 6),
             GateDesc("poseidon")]
    groups = [(0, 5), (5, 10), (10, 12), (12, 13), (13, 14)]
    selector_indices = [s for s, (a, b) in enumerate(groups) for _ in range(a, b)]
    return CircuitDesc(degree_bits, NUM_WIRES, NUM_ROUTED_WIRES, len(groups) + 2, gates, selector_indices, groups, 3, 2)


def trace_selector_columns(desc, n: int, rng):
    """One gate index per row, written into the selector column of its group (UNUSED_SELECTOR = 2^32 - 1 elsewhere)."""
    import numpy as np

    row_gate = rng.integers(0, len(desc.gates), n)
    cols = np.full((len(desc.groups), n), (1 << 32) - 1, dtype=np.uint64)
    for s, (a, b) in enumerate(desc.groups):
        m = (row_gate >= a) & (row_gate < b)
        cols[s, m] = row_gate[m].astype(np.uint64)
    return cols


class ProverTrace:
    """``nthreads`` independent provers on the current device, each replaying whole proofs."""

    def __init__(self, degrees=LEAF_PROOF_DEGREES, hash_kind: int = 1, nthreads: int = 8, seed: int = 0x7ACE,
                 native: bool = True):
        import numpy as np

        from . import fri as GF
        from . import plonky2 as P2

        self.np, self.GF, self.P2 = np, GF, P2
        self.degrees, self.hash_kind, self.nthreads = tuple(degrees), hash_kind, nthreads
        self.native = native    # mp2gpu_prove (csrc/prover.cpp) instead of the Python mirror's call sequence
        self.stage_log = None   # set to a list to collect (degree, [(stage, ms), ...]) per prove() (Python mirror only)
        rng = np.random.default_rng(seed)
        self.circuits = {}
        for d in set(degrees):
            n = 1 << d
            desc = trace_circuit_desc(d)
            cs = rng.integers(0, 1 << 62, (desc.num_constants + NUM_ROUTED_WIRES, n), dtype=np.uint64)
            cs[:desc.num_selectors] = trace_selector_columns(desc, n, rng)
            batch = P2.PolynomialBatch.from_values(cs, RATE_BITS, False, CAP_HEIGHT, hash_kind=hash_kind, keep_on_device=True,
                                                   fetch_leaves=False)
            self.circuits[d] = (desc, batch)
        # per-thread synthetic witnesses (host memory; reused across proofs, the work does not depend on the values)
        def pinned(shape):
            a = P2.pinned_empty(shape)
            a[...] = rng.integers(0, 1 << 62, shape, dtype=np.uint64)
            return a

        self.inputs = [{d: (pinned((NUM_WIRES, 1 << d)), pinned((ZS_PP_COLS, 1 << d))) for d in set(degrees)}
                       for _ in range(nthreads)]

    def prove(self, degree_bits: int, wires, zs_pp):
        np, GF, P2 = self.np, self.GF, self.P2
        from .quotient import compute_quotient_polys

        if self.native and self.stage_log is None:
            from .prover import prove_native

            desc, cs = self.circuits[degree_bits]
            return prove_native(desc, cs, [5, 6, 7, 8], wires, [], [1, 2, 3, 4], GF.FriConfig(), hash_kind=self.hash_kind)

        import time

        desc, cs = self.circuits[degree_bits]
        kind = self.hash_kind
        marks = [("start", time.perf_counter())]
        mark = lambda name: marks.append((name, time.perf_counter()))
        ch = GF.Challenger(kind)
        ch.observe_cap(cs.merkle_tree.cap)
        commit = lambda cols: P2.PolynomialBatch.from_values(cols, RATE_BITS, False, CAP_HEIGHT, hash_kind=kind,
                                                             keep_on_device=True, fetch_leaves=False, fetch_coeffs=False,
                                                             fetch_digests=False)
        b_w = commit(wires)
        mark("commit_wires")
        ch.observe_cap(b_w.merkle_tree.cap)
        betas, gammas = ch.get_n_challenges(2), ch.get_n_challenges(2)
        b_z = commit(zs_pp)
        mark("commit_zs")
        ch.observe_cap(b_z.merkle_tree.cap)
        alphas = ch.get_n_challenges(2)
        b_q = compute_quotient_polys(desc, cs, b_w, b_z, betas, gammas, alphas, [1, 2, 3, 4], RATE_BITS, CAP_HEIGHT,
                                     hash_kind=kind, fetch_digests=False, fetch_chunks=False)
        mark("quotient")
        ch.observe_cap(b_q.merkle_tree.cap)
        zeta = ch.get_extension_challenge()
        g = pow(7, (P2.ORDER - 1) >> degree_bits, P2.ORDER)  # any point: the replay only needs the work
        gz = np.array([int(zeta[0]) * g % P2.ORDER, int(zeta[1]) * g % P2.ORDER], dtype=np.uint64)
        oracles = [cs, b_w, b_z, b_q]
        batches = [GF.FriBatchInfo(zeta, [(o, p) for o, b in enumerate(oracles) for p in range(b.num_polys)]),
                   GF.FriBatchInfo(gz, [(2, 0), (2, 1)])]
        openings = GF.open_batches(batches, oracles)
        mark("openings")
        for v in openings:
            ch.observe_extension_elements(v)
        proof = GF.prove_openings(batches, oracles, ch, GF.FriConfig().fri_params(degree_bits))
        mark("fri")
        for b in (b_w, b_z, b_q):
            b.free()
        mark("free")
        if self.stage_log is not None:
            self.stage_log.append((degree_bits, [(b[0], (b[1] - a[1]) * 1e3) for a, b in zip(marks, marks[1:])]))
        return proof

    def _start_workers(self) -> None:
        """Persistent prover threads: the library keeps one stream set per calling thread and the stream-ordered pool
        hands freed blocks back to the stream that freed them, so a prover must stay on its thread (fresh threads per
        run made every run re-grow the pool: 10x swings in the measured rate)."""
        import queue
        import threading

        from . import plonky2 as P2

        self._jobs, self._done = queue.Queue(), queue.Queue()
        device = self.torch_device()

        def worker(t):
            try:
                P2.init(device)
            except Exception as e:  # noqa: BLE001
                self._done.put(e)
                return
            while True:
                job = self._jobs.get()
                if job is None:
                    return
                try:
                    for d in self.degrees:
                        w, z = self.inputs[t][d]
                        self.prove(d, w, z)
                    self._done.put(None)
                except Exception as e:  # noqa: BLE001
                    self._done.put(e)

        self._threads = [threading.Thread(target=worker, args=(t,), daemon=True) for t in range(self.nthreads)]
        for th in self._threads:
            th.start()

    def run(self, nproofs: int) -> None:
        """``nproofs`` whole proofs, shared out over the prover threads; returns when all are done."""
        import sys

        if not getattr(self, "_threads", None):
            self._start_workers()
        # the provers spend their time inside ctypes calls (GIL released); a thread that returns from one must not
        # wait the default 5 ms switch interval for the interpreter
        old = sys.getswitchinterval()
        sys.setswitchinterval(5e-5)
        try:
            for _ in range(nproofs):
                self._jobs.put(1)
            errors = [e for e in (self._done.get() for _ in range(nproofs)) if e is not None]
        finally:
            sys.setswitchinterval(old)
        if errors:
            raise errors[0]

    @staticmethod
    def torch_device() -> int:
        import torch

        return torch.cuda.current_device()

    def free(self) -> None:
        for _ in getattr(self, "_threads", []):
            self._jobs.put(None)
        for th in getattr(self, "_threads", []):
            th.join(timeout=10)
        self._threads = []
        for _, b in self.circuits.values():
            b.free()
