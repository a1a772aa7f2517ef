"""Wall-clock of the device FRI prover at the shapes of one prove() (degree 2^k, oracles 85/135/20/16 columns,
standard_recursion_config): openings, prove_openings' quotient, commit phase + PoW + 28 query rounds."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mapreduce_plonky2_b200 as G  # noqa: E402
from mapreduce_plonky2_b200 import fri as GF  # noqa: E402

P = 0xFFFFFFFF00000001


def main():
    kind = int(os.environ.get("KIND", "1"))
    G.init(0)
    rng = np.random.default_rng(7)
    for degree_bits in (12, 13, 14):
        n = 1 << degree_bits
        widths = (85, 135, 20, 16)
        t0 = time.perf_counter()
        oracles = [G.PolynomialBatch.from_coeffs(list(rng.integers(0, P, (w, n), dtype=np.uint64)), 3, False, 4,
                                                 hash_kind=kind, keep_on_device=True, fetch_leaves=False) for w in widths]
        t_commit = time.perf_counter() - t0
        zeta, gzeta = rng.integers(0, P, 2, dtype=np.uint64), rng.integers(0, P, 2, dtype=np.uint64)
        batches = [GF.FriBatchInfo(zeta, [(o, p) for o, w in enumerate(widths) for p in range(w)]),
                   GF.FriBatchInfo(gzeta, [(2, 0), (2, 1)])]
        params = GF.FriConfig().fri_params(degree_bits)
        best = None
        for it in range(4):
            ch = GF.Challenger(kind)
            for o in oracles:
                ch.observe_cap(o.merkle_tree.cap)
            t0 = time.perf_counter()
            openings = GF.open_batches(batches, oracles)
            t1 = time.perf_counter()
            for v in openings:
                ch.observe_extension_elements(v)
            t2 = time.perf_counter()
            alpha = ch.get_extension_challenge()
            phase = G.FriCommitPhase.from_openings(oracles, [(b.point, b.polynomials) for b in batches], alpha, 4, kind)
            t3 = time.perf_counter()
            proof = GF.fri_proof(oracles, phase, ch, params)
            t4 = time.perf_counter()
            phase.free()
            cur = (t1 - t0, t2 - t1, t3 - t2, t4 - t3)
            best = cur if best is None or sum(cur) < sum(best) else best
        print("degree 2^%d kind %d: 4 commitments (host in/out) %.1f ms | openings %.2f ms, transcript of openings %.2f ms, "
              "prove_openings quotient %.2f ms, fri_proof (commit phase + PoW + %d query rounds) %.2f ms; pow_witness %d"
              % (degree_bits, kind, t_commit * 1e3, best[0] * 1e3, best[1] * 1e3, best[2] * 1e3, len(proof.query_round_proofs),
                 best[3] * 1e3, proof.pow_witness), flush=True)
        for o in oracles:
            o.free()


if __name__ == "__main__":
    main()
