//! Golden-vector dump from the real plonky2 fork (see Cargo.toml).  Output schema = tests/golden/commit_small.json
//! ("cases": cols / coeffs / leaves / digests / cap as 16-digit hex) plus a second file of hashing, tree,
//! serialization and proof vectors.  What each vector pins:
//!
//!   reference_commit.json
//!     * PolynomialBatch::from_values 4 x 2^3, rate_bits 1, cap_height 1, both hashers;
//!       more small shapes incl. from_coeffs and non-canonical inputs
//!     * 2^14 x 135, rate_bits 3, cap_height 4 on SplitMix64(0x6d7032) inputs: cap + digest checksum
//!       (BASELINE config 1; mp2-common/src/lib.rs:37-47)
//!   reference_kats.json
//!     * Poseidon2 permutation / hash_no_pad / two_to_one / hash_pad(&[]) vectors (is poseidon2_plonky2 the
//!       Horizen-Labs t = 12 instance with plonky2's sponge?  SURVEY.md Appendix C.1)
//!     * MerkleTree::new on the circuit-set shape (recursion-framework/src/universal_verifier_gadget/
//!       circuit_set.rs:173-191: 4-element digests padded with vec![F::ZERO], cap_height 0), prove(i) for every i,
//!       and Buffer::write_merkle_tree bytes (mp2-common/src/serialization/circuit_data_serialization.rs:74-89)
//!     * a tiny circuit proved with standard_recursion_config: bincode(ProofWithPublicInputs) bytes, the three
//!       caps, the FRI caps, final polynomial and pow_witness (mp2-common/src/proof.rs:41-57)
use std::{fs, path::PathBuf};

use anyhow::Result;
use plonky2::field::extension::Extendable;
use plonky2::field::goldilocks_field::GoldilocksField;
use plonky2::field::polynomial::{PolynomialCoeffs, PolynomialValues};
use plonky2::field::types::{Field, PrimeField64};
use plonky2::fri::oracle::PolynomialBatch;
use plonky2::hash::hash_types::{HashOut, RichField};
use plonky2::hash::merkle_tree::MerkleTree;
use plonky2::hash::poseidon::{Poseidon, PoseidonHash};
use plonky2::iop::witness::{PartialWitness, WitnessWrite};
use plonky2::plonk::circuit_builder::CircuitBuilder;
use plonky2::plonk::circuit_data::CircuitConfig;
use plonky2::plonk::config::{GenericConfig, Hasher, PoseidonGoldilocksConfig};
use plonky2::util::serialization::Write;
use plonky2::util::timing::TimingTree;
use poseidon2_plonky2::poseidon2_goldilock::Poseidon2GoldilocksConfig;
use poseidon2_plonky2::poseidon2_hash::{Poseidon2, Poseidon2Hash};
use serde_json::{json, Value};

type F = GoldilocksField;
const D: usize = 2;

fn hx(x: F) -> String {
    format!("{:016x}", x.to_canonical_u64())
}
fn hxv(v: &[F]) -> Vec<String> {
    v.iter().map(|x| hx(*x)).collect()
}
fn hash_hex(h: &HashOut<F>) -> Vec<String> {
    hxv(&h.elements)
}
fn bytes_hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{:02x}", x)).collect()
}

/// tests/util.py::splitmix64 + rejection below p (same stream as the Python tests)
struct SplitMix(u64);
impl SplitMix {
    fn next(&mut self) -> u64 {
        self.0 = self.0.wrapping_add(0x9E3779B97F4A7C15);
        let mut z = self.0;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
        z ^ (z >> 31)
    }
}

fn commit_case<C: GenericConfig<D, F = F>>(
    kind: u32,
    cols: Vec<Vec<u64>>,
    rate_bits: usize,
    cap_height: usize,
    from_coeffs: bool,
    full: bool,
) -> Value {
    let mut timing = TimingTree::default();
    let n = cols[0].len();
    let as_f = |c: &Vec<u64>| c.iter().map(|&x| F::from_noncanonical_u64(x)).collect::<Vec<F>>();
    let batch: PolynomialBatch<F, C, D> = if from_coeffs {
        PolynomialBatch::from_coeffs(
            cols.iter().map(|c| PolynomialCoeffs::new(as_f(c))).collect(),
            rate_bits,
            false,
            cap_height,
            &mut timing,
            None,
        )
    } else {
        PolynomialBatch::from_values(
            cols.iter().map(|c| PolynomialValues::new(as_f(c))).collect(),
            rate_bits,
            false,
            cap_height,
            &mut timing,
            None,
        )
    };
    let hash_of = |h: &<C::Hasher as Hasher<F>>::Hash| -> Vec<String> {
        // HashOut<F> for both Goldilocks hashers; go through its bytes to stay generic
        let b = plonky2::plonk::config::GenericHashOut::<F>::to_vec(h);
        hxv(&b)
    };
    let mut v = json!({
        "hash_kind": kind, "ncols": cols.len(), "log_n": n.trailing_zeros(), "rate_bits": rate_bits,
        "cap_height": cap_height, "from_coeffs": from_coeffs,
        "cap": batch.merkle_tree.cap.0.iter().map(|h| hash_of(h)).collect::<Vec<_>>(),
    });
    if full {
        v["cols"] = json!(cols.iter().map(|c| c.iter().map(|x| format!("{:016x}", x)).collect::<Vec<_>>()).collect::<Vec<_>>());
        v["coeffs"] = json!(batch.polynomials.iter().map(|p| hxv(&p.coeffs)).collect::<Vec<_>>());
        v["leaves"] = json!(batch.merkle_tree.leaves.iter().map(|l| hxv(l)).collect::<Vec<_>>());
        v["digests"] = json!(batch.merkle_tree.digests.iter().map(|h| hash_of(h)).collect::<Vec<_>>());
    } else {
        // big case: inputs are re-derived from the seed by the test; pin the cap and an xor over the digests
        let mut x = [0u64; 4];
        for h in &batch.merkle_tree.digests {
            let e = plonky2::plonk::config::GenericHashOut::<F>::to_vec(h);
            for i in 0..4 {
                x[i] ^= e[i].to_canonical_u64();
            }
        }
        v["digests_xor"] = json!(x.iter().map(|y| format!("{:016x}", y)).collect::<Vec<_>>());
        v["seed"] = json!("6d7032");
    }
    v
}

fn commit_cases() -> Value {
    let mut cases = vec![];
    let mut rng = SplitMix(0x6d7032);
    // (ncols, log_n, rate_bits, cap_height, from_coeffs) -- first entry is SURVEY 8(c)'s 4 x 2^3, r = 1, cap = 1
    let shapes = [(4, 3, 1, 1, false), (5, 3, 1, 1, false), (9, 2, 2, 0, false), (3, 1, 3, 4, false), (17, 2, 1, 2, true), (12, 3, 3, 4, false)];
    for kind in 0..2u32 {
        for &(c, ln, r, cap, fc) in shapes.iter() {
            let mut cols: Vec<Vec<u64>> = (0..c).map(|_| (0..1usize << ln).map(|_| loop {
                let x = rng.next();
                if x < F::ORDER { break x; }
            }).collect()).collect();
            if c == 9 {
                cols[0][0] = F::ORDER + 5; // non-canonical inputs must be accepted
                cols[1][1] = u64::MAX;
            }
            cases.push(if kind == 0 {
                commit_case::<PoseidonGoldilocksConfig>(kind, cols, r, cap, fc, true)
            } else {
                commit_case::<Poseidon2GoldilocksConfig>(kind, cols, r, cap, fc, true)
            });
        }
        // BASELINE config 1: the raw SplitMix64 stream (tests/util.py::splitmix64(0x6d7032, 135 << 14), column-major;
        // values >= p are legal non-canonical inputs on both sides)
        let mut big = SplitMix(0x6d7032);
        let cols: Vec<Vec<u64>> = (0..135).map(|_| (0..1usize << 14).map(|_| big.next()).collect()).collect();
        cases.push(if kind == 0 {
            commit_case::<PoseidonGoldilocksConfig>(kind, cols, 3, 4, false, false)
        } else {
            commit_case::<Poseidon2GoldilocksConfig>(kind, cols, 3, 4, false, false)
        });
    }
    json!({"generator": "tools/golden_dump (plonky2 fork rev 22c42f64)", "cases": cases})
}

fn hasher_vectors<H: Hasher<F, Hash = HashOut<F>>>(perm: impl Fn([F; 12]) -> [F; 12]) -> Value {
    let iota: [F; 12] = core::array::from_fn(|i| F::from_canonical_u64(i as u64));
    let neg1 = [F::NEG_ONE; 12];
    let inputs: Vec<Vec<F>> = vec![vec![], (0..1).map(F::from_canonical_u64).collect(), (0..4).map(F::from_canonical_u64).collect(),
        (0..5).map(F::from_canonical_u64).collect(), (0..8).map(F::from_canonical_u64).collect(), (0..9).map(F::from_canonical_u64).collect(),
        (0..135).map(F::from_canonical_u64).collect()];
    json!({
        "perm_zeros": hxv(&perm([F::ZERO; 12])), "perm_iota": hxv(&perm(iota)), "perm_neg_one": hxv(&perm(neg1)),
        "hash_no_pad": inputs.iter().map(|x| json!({"len": x.len(), "out": hash_hex(&H::hash_no_pad(x))})).collect::<Vec<_>>(),
        "hash_or_noop": inputs.iter().map(|x| json!({"len": x.len(), "out": hash_hex(&H::hash_or_noop(x))})).collect::<Vec<_>>(),
        "hash_pad_empty": hash_hex(&H::hash_pad(&[])),
        "two_to_one_01": hash_hex(&H::two_to_one(H::hash_no_pad(&[F::ZERO]), H::hash_no_pad(&[F::ONE]))),
    })
}

fn circuit_set_tree<H: Hasher<F, Hash = HashOut<F>>>() -> Value {
    // 42 digests (as at circuit_set.rs:173-191) padded to 64 leaves with vec![F::ZERO], cap_height 0
    let mut leaves: Vec<Vec<F>> = (0..42u64).map(|i| H::hash_no_pad(&[F::from_canonical_u64(i)]).elements.to_vec()).collect();
    leaves.resize(64, vec![F::ZERO]);
    let tree = MerkleTree::<F, H>::new(leaves.clone(), 0);
    let mut bytes = Vec::new();
    bytes.write_merkle_tree(&tree).unwrap();
    json!({
        "leaves": leaves.iter().map(|l| hxv(l)).collect::<Vec<_>>(),
        "digests": tree.digests.iter().map(hash_hex).collect::<Vec<_>>(),
        "cap": tree.cap.0.iter().map(hash_hex).collect::<Vec<_>>(),
        "proofs": (0..64).map(|i| tree.prove(i).siblings.iter().map(hash_hex).collect::<Vec<_>>()).collect::<Vec<_>>(),
        "write_merkle_tree": bytes_hex(&bytes),
    })
}

fn tiny_proof<C: GenericConfig<D, F = F>>() -> Result<Value>
where
    F: RichField + Extendable<D>,
{
    // x^2 * y public, enough gates to reach a few hundred rows; standard_recursion_config (mp2-common/src/lib.rs:45-47)
    let mut b = CircuitBuilder::<F, D>::new(CircuitConfig::standard_recursion_config());
    let x = b.add_virtual_target();
    let y = b.add_virtual_target();
    let mut acc = b.mul(x, x);
    for _ in 0..300 {
        acc = b.mul(acc, y);
        acc = b.add(acc, x);
    }
    b.register_public_input(acc);
    let data = b.build::<C>();
    let mut pw = PartialWitness::new();
    pw.set_target(x, F::from_canonical_u64(3));
    pw.set_target(y, F::from_canonical_u64(5));
    let proof = data.prove(pw)?;
    data.verify(proof.clone())?;
    let fri = &proof.proof.opening_proof;
    Ok(json!({
        "degree_bits": data.common.degree_bits(),
        "bincode_proof_with_public_inputs": bytes_hex(&bincode::serialize(&proof)?),
        "verifier_only_to_bytes": bytes_hex(&data.verifier_only.to_bytes().unwrap()),
        "public_inputs": hxv(&proof.public_inputs),
        "pow_witness": hx(fri.pow_witness),
        "num_fri_layers": fri.commit_phase_merkle_caps.len(),
        "final_poly_len": fri.final_poly.len(),
    }))
}

fn main() -> Result<()> {
    let out = PathBuf::from(std::env::args().nth(1).unwrap_or_else(|| "../../tests/golden".into()));
    fs::write(out.join("reference_commit.json"), serde_json::to_string(&commit_cases())?)?;
    let kats = json!({
        "generator": "tools/golden_dump (plonky2 fork rev 22c42f64)",
        "poseidon": hasher_vectors::<PoseidonHash>(|s| <F as Poseidon>::poseidon(s)),
        "poseidon2": hasher_vectors::<Poseidon2Hash>(|s| <F as Poseidon2>::poseidon2(s)),
        "circuit_set_tree": {"poseidon": circuit_set_tree::<PoseidonHash>(), "poseidon2": circuit_set_tree::<Poseidon2Hash>()},
        "tiny_proof": {"poseidon": tiny_proof::<PoseidonGoldilocksConfig>()?, "poseidon2": tiny_proof::<Poseidon2GoldilocksConfig>()?},
    });
    fs::write(out.join("reference_kats.json"), serde_json::to_string(&kats)?)?;
    println!("wrote reference_commit.json and reference_kats.json to {}", out.display());
    Ok(())
}
