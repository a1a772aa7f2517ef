"""mp2gpu -- B200-native (sm_100a) polynomial-batch commitment for plonky2 / mapreduce-plonky2.

The package holds only what the hot path needs: ``csrc/`` (hand-written CUDA kernels + the C ABI of
``include/mp2gpu.h``), ``plonky2.py`` (host-side mirror of the plonky2 surface the reference calls),
``fri.py`` (challenger, ``prove_openings`` / ``fri_proof`` glue over the device primitives), ``device.py``
(device-pointer stages for resident data) and ``sharded.py`` (multi-GPU driver).
"""
from ._lib import LIB_PATH, Mp2GpuError  # noqa: F401
from .plonky2 import (POSEIDON, POSEIDON2, FriCommitPhase, fri_committed_trees, fri_proof_of_work, Communicator, MerkleCap, MerkleProof, MerkleTree, PolynomialBatch,  # noqa: F401
                      circuit_digest, device_count, hash_no_pad, hash_no_pad_batch, hash_or_noop, hash_pad, init, launch_count,
                      permute, reverse_bits, two_to_one, two_to_one_batch, verify_merkle_proof_to_cap, read_merkle_tree, write_merkle_tree,
                      read_polynomial_batch, write_polynomial_batch)

from .fri import (Challenger, FriBatchInfo, FriConfig, FriParams, FriProof, fri_proof, open_batches,  # noqa: F401,E402
                  prove_openings)

from . import wire  # noqa: F401,E402  (bincode mirrors of FriProof / ProofWithPublicInputs / ProofWithVK)
from . import quotient  # noqa: F401,E402  (compute_quotient_polys on device-resident batches)

__version__ = "0.2.0"
