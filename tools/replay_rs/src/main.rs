//! Replays a kit made by tools/make_replay_kit.py through the real arkworks Marlin and diffs bytes.  See Cargo.toml.
//! (BLS12-381 only: ark-bn254 is not a dependency of the reference.)
use ark_bls12_381::{Bls12_381, Fr, G1Affine, G2Affine};
use ark_ff::{to_bytes, Field, PrimeField};
use ark_marlin::{Marlin, SimpleHashFiatShamirRng};
use ark_poly::univariate::DensePolynomial;
use ark_poly_commit::{kzg10::UniversalParams, marlin_pc::MarlinKZG10, sonic_pc::SonicKZG10, PolynomialCommitment};
use ark_relations::{lc, r1cs::{ConstraintSynthesizer, ConstraintSystemRef, SynthesisError}};
use ark_serialize::{CanonicalDeserialize, CanonicalSerialize};
use blake2::Blake2s;
use rand::SeedableRng;
use rand_chacha::{ChaCha12Rng, ChaChaRng};
use std::collections::BTreeMap;
use std::io::Read;

type FS = SimpleHashFiatShamirRng<Blake2s, ChaChaRng>;

/// The reference bench's circuit (benches/bench.rs:25-67), with the witness values of the kit.
#[derive(Copy, Clone)]
struct DummyCircuit { a: Option<Fr>, b: Option<Fr>, num_variables: usize, num_constraints: usize }
impl ConstraintSynthesizer<Fr> for DummyCircuit {
    fn generate_constraints(self, cs: ConstraintSystemRef<Fr>) -> Result<(), SynthesisError> {
        let a = cs.new_witness_variable(|| self.a.ok_or(SynthesisError::AssignmentMissing))?;
        let b = cs.new_witness_variable(|| self.b.ok_or(SynthesisError::AssignmentMissing))?;
        let c = cs.new_input_variable(|| Ok(self.a.ok_or(SynthesisError::AssignmentMissing)? * self.b.ok_or(SynthesisError::AssignmentMissing)?))?;
        for _ in 0..(self.num_variables - 3) {
            let _ = cs.new_witness_variable(|| self.a.ok_or(SynthesisError::AssignmentMissing))?;
        }
        for _ in 0..self.num_constraints - 1 {
            cs.enforce_constraint(lc!() + a, lc!() + b, lc!() + c)?;
        }
        cs.enforce_constraint(lc!(), lc!(), lc!())?;
        Ok(())
    }
}

fn read_u64(r: &mut impl Read) -> u64 { let mut b = [0u8; 8]; r.read_exact(&mut b).unwrap(); u64::from_le_bytes(b) }

/// marlin_b200/srsfile.py layout: tag, curve id, then the public fields of kzg10::UniversalParams, each `serialize_uncompressed`.
fn load_srs(path: &str) -> UniversalParams<Bls12_381> {
    let mut f = std::io::BufReader::new(std::fs::File::open(path).expect("srs file"));
    let mut tag = [0u8; 8];
    f.read_exact(&mut tag).unwrap();
    assert_eq!(&tag, b"B2MSRS01");
    assert_eq!(read_u64(&mut f), 0, "curve id 0 = BLS12-381");
    let n = read_u64(&mut f) as usize;
    let powers_of_g: Vec<G1Affine> = (0..n).map(|_| G1Affine::deserialize_unchecked(&mut f).unwrap()).collect();
    let mut powers_of_gamma_g = BTreeMap::new();
    for _ in 0..read_u64(&mut f) { let k = read_u64(&mut f) as usize; powers_of_gamma_g.insert(k, G1Affine::deserialize_unchecked(&mut f).unwrap()); }
    let h = G2Affine::deserialize_unchecked(&mut f).unwrap();
    let beta_h = G2Affine::deserialize_unchecked(&mut f).unwrap();
    let mut neg_powers_of_h = BTreeMap::new();
    for _ in 0..read_u64(&mut f) { let k = read_u64(&mut f) as usize; neg_powers_of_h.insert(k, G2Affine::deserialize_unchecked(&mut f).unwrap()); }
    UniversalParams { powers_of_g, powers_of_gamma_g, h, beta_h, neg_powers_of_h, prepared_h: h.into(), prepared_beta_h: beta_h.into() }
}

fn fr_from_dec(s: &str) -> Fr { Fr::from_str(s).ok().expect("decimal field element") }

fn replay<PC>(dir: &str, meta: &serde_json::Value, srs: &PC::UniversalParams) -> bool
where PC: PolynomialCommitment<Fr, DensePolynomial<Fr>> {
    let scheme = meta["pc"].as_str().unwrap();
    let circ = DummyCircuit { a: Some(fr_from_dec(meta["a"].as_str().unwrap())), b: Some(fr_from_dec(meta["b"].as_str().unwrap())),
                              num_variables: meta["num_variables"].as_u64().unwrap() as usize, num_constraints: meta["num_constraints"].as_u64().unwrap() as usize };
    let (pk, vk) = Marlin::<Fr, PC, FS>::index(srs, circ).expect("index");
    let want_vk = std::fs::read(format!("{}/{}_index_vk_tobytes.bin", dir, scheme)).unwrap();
    let got_vk = to_bytes![vk].unwrap();
    let mut seed = [0u8; 32];
    seed.copy_from_slice(&hex::decode(meta["zk_seed_hex"].as_str().unwrap()).unwrap());
    let mut zk = ChaCha12Rng::from_seed(seed);
    let proof = Marlin::<Fr, PC, FS>::prove(&pk, circ, &mut zk).expect("prove");
    let mut got_proof = Vec::new();
    proof.serialize(&mut got_proof).unwrap();
    let want_proof = std::fs::read(format!("{}/{}_proof.bin", dir, scheme)).unwrap();
    let c = circ.a.unwrap() * circ.b.unwrap();
    let ok_verify = Marlin::<Fr, PC, FS>::verify(&vk, &[c], &proof, &mut ChaCha12Rng::from_seed([7u8; 32])).expect("verify");
    let want_pos = meta["zk_word_pos_after"][scheme].as_u64().unwrap() as u128;
    println!("{}: index_vk {} | proof {} | rng position {} | verify {}", scheme,
             if got_vk == want_vk { "MATCH" } else { "DIFFER" }, if got_proof == want_proof { "MATCH" } else { "DIFFER" },
             if zk.get_word_pos() == want_pos { "MATCH" } else { "DIFFER" }, ok_verify);
    // the GPU-made proof must also be accepted by the real verifier
    let gpu_proof = ark_marlin::Proof::<Fr, PC>::deserialize(&want_proof[..]).expect("deserialize the kit's proof");
    let ok_gpu = Marlin::<Fr, PC, FS>::verify(&vk, &[c], &gpu_proof, &mut ChaCha12Rng::from_seed([7u8; 32])).unwrap_or(false);
    println!("{}: the kit's proof is accepted by ark-marlin: {}", scheme, ok_gpu);
    got_vk == want_vk && got_proof == want_proof && zk.get_word_pos() == want_pos && ok_verify && ok_gpu
}

fn main() {
    let dir = std::env::args().nth(1).expect("usage: replay_rs <kit directory>");
    let meta: serde_json::Value = serde_json::from_slice(&std::fs::read(format!("{}/meta.json", dir)).unwrap()).unwrap();
    let srs = load_srs(&format!("{}/srs.bin", dir));
    let a = replay::<MarlinKZG10<Bls12_381, DensePolynomial<Fr>>>(&dir, &meta, &srs);
    let b = replay::<SonicKZG10<Bls12_381, DensePolynomial<Fr>>>(&dir, &meta, &srs);
    std::process::exit(if a && b { 0 } else { 1 });
}
