//! Safe wrappers over libb2m.so for ark-marlin 0.3 on BLS12-381 (include/b2m.h).
//!
//! Seams (reference file:line):
//!  * `msm`            replaces `VariableBaseMSM::multi_scalar_mul` inside `KZG10::commit` / `open`
//!                     [U ark-poly-commit kzg10], called from `src/lib.rs:125,172,193,213,292`;
//!  * `fft_in_place`   replaces `EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place`
//!                     (`src/ahp/prover.rs:321-686`);
//!  * `IndexProverKey::new` + `prove` replace `Marlin::index` (`src/lib.rs:100-148`) and `Marlin::prove`
//!                     (`src/lib.rs:151-311`); the proof comes back as ark-serialize bytes and is deserialised into the
//!                     reference's own `Proof` type, so `Marlin::verify` (`src/lib.rs:315-433`) runs unchanged on the CPU.
//!
//!  * `pc::B200MarlinKZG10` / `pc::B200SonicKZG10` implement `PolynomialCommitment` (the generic parameter `PC` of
//!                     `Marlin<F, PC, FS>`, `src/lib.rs:64-71`): `setup` / `check*` delegate to ark-poly-commit, `trim` / `commit` /
//!                     `open*` run on the GPU, so `Marlin::<Fr, B200MarlinKZG10, FS>::{index, prove, verify}` is the stock code.
//!
//! This crate has NOT been compiled (the build image has no Rust toolchain); the C side it binds is tested through the
//! Python ctypes twin of these declarations (marlin_b200/_lib.py, tests/).
pub mod ffi;
pub mod pc;

use ark_bls12_381::{Bls12_381, Fq, Fr, G1Affine, G1Projective};
use ark_ec::{AffineCurve, ProjectiveCurve};
use ark_ff::{BigInteger256, BigInteger384, Field, One, PrimeField, Zero};
use ark_relations::r1cs::{ConstraintMatrices, ConstraintSynthesizer, ConstraintSystem, OptimizationGoal, SynthesisMode};
use ark_relations::lc;
use ark_serialize::CanonicalDeserialize;
use ark_std::vec::Vec;
use rand_chacha::ChaCha12Rng;
use std::ffi::CStr;
use std::os::raw::c_int;

#[derive(Debug)]
pub enum Error {
    /// `ark_marlin::Error::IndexTooLarge`
    IndexTooLarge,
    /// `ark_marlin::ahp::Error::{InstanceDoesNotMatchIndex, InvalidPublicInputLength, NonSquareMatrix}`
    InstanceDoesNotMatchIndex,
    InvalidPublicInputLength,
    NonSquareMatrix,
    /// `SynthesisError::PolynomialDegreeTooLarge` / ark-poly-commit degree errors
    DegreeTooLarge,
    /// `ark_poly_commit::Error::MissingRng`
    MissingRng,
    Synthesis(ark_relations::r1cs::SynthesisError),
    Serialization(ark_serialize::SerializationError),
    /// CUDA / NCCL / argument errors, with the library's message
    Device(c_int, String),
}

pub(crate) fn check(code: c_int) -> Result<(), Error> {
    use ffi::*;
    match code {
        B2M_OK => Ok(()),
        B2M_ERR_INDEX_TOO_LARGE => Err(Error::IndexTooLarge),
        B2M_ERR_INSTANCE_MISMATCH => Err(Error::InstanceDoesNotMatchIndex),
        B2M_ERR_INVALID_PUBLIC_INPUT_LEN => Err(Error::InvalidPublicInputLength),
        B2M_ERR_NON_SQUARE => Err(Error::NonSquareMatrix),
        B2M_ERR_DEGREE_TOO_LARGE => Err(Error::DegreeTooLarge),
        B2M_ERR_MISSING_RNG => Err(Error::MissingRng),
        c => {
            let msg = unsafe { CStr::from_ptr(b2m_last_error()) }.to_string_lossy().into_owned();
            Err(Error::Device(c, msg))
        }
    }
}

// ---- marshalling: the library uses ark-ff's own Montgomery limbs (INTEGRATION.md "Marshalling rules") --------------------
fn fq_limbs(x: &Fq) -> [u64; 6] {
    (x.0).0
}
fn fq_from_limbs(l: &[u64]) -> Fq {
    let mut a = [0u64; 6];
    a.copy_from_slice(&l[..6]);
    Fq::new(BigInteger384(a)) // `new` takes the Montgomery representation as is
}
pub(crate) fn g1_limbs(p: &G1Affine, out: &mut Vec<u64>) {
    if p.infinity {
        out.extend_from_slice(&[0u64; 12]);
    } else {
        out.extend_from_slice(&fq_limbs(&p.x));
        out.extend_from_slice(&fq_limbs(&p.y));
    }
}
pub(crate) fn g1_from_limbs(l: &[u64], is_inf: bool) -> G1Affine {
    if is_inf || l[..12].iter().all(|w| *w == 0) {
        G1Affine::zero()
    } else {
        G1Affine::new(fq_from_limbs(&l[..6]), fq_from_limbs(&l[6..12]), false)
    }
}
/// Montgomery limbs of field elements (polynomial coefficients, witness values): `fe.0.0` verbatim.
pub fn fr_mont_limbs(v: &[Fr]) -> Vec<u64> {
    v.iter().flat_map(|x| (x.0).0.iter().copied().collect::<Vec<_>>()).collect()
}
/// Canonical limbs of MSM scalars: `into_repr()`, what `VariableBaseMSM::multi_scalar_mul` receives.
pub fn fr_canonical_limbs(v: &[Fr]) -> Vec<u64> {
    v.iter().flat_map(|x| x.into_repr().0.iter().copied().collect::<Vec<_>>()).collect()
}

/// One GPU (one CUDA device + streams).  Not `Sync`: use one context per thread.
pub struct Context {
    pub(crate) raw: *mut ffi::b2m_ctx,
}
impl Context {
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::b2m_ctx_create(device, &mut raw) })?;
        Ok(Context { raw })
    }
    /// `domain.fft_in_place` / `ifft_in_place` / `coset_*` for a radix-2 domain of size `coeffs.len()` (a power of two).
    pub fn fft_in_place(&self, coeffs: &mut Vec<Fr>, inverse: bool, coset: bool) -> Result<(), Error> {
        assert!(coeffs.len().is_power_of_two());
        let mut limbs = fr_mont_limbs(coeffs);
        let log_n = coeffs.len().trailing_zeros();
        check(unsafe { ffi::b2m_ntt(self.raw, ffi::B2M_CURVE_BLS12_381, limbs.as_mut_ptr(), log_n, inverse as c_int, coset as c_int) })?;
        for (c, l) in coeffs.iter_mut().zip(limbs.chunks(4)) {
            let mut a = [0u64; 4];
            a.copy_from_slice(l);
            *c = Fr::new(BigInteger256(a));
        }
        Ok(())
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::b2m_ctx_destroy(self.raw) }
    }
}

/// The committer key on the device: `powers_of_g` (window tables are built here, once -- this is `PC::trim`) and the
/// `powers_of_gamma_g` entries the PC variant will use, each with its index in the full gamma-power list.
pub struct Srs<'c> {
    raw: *mut ffi::b2m_srs,
    _ctx: &'c Context,
}
impl<'c> Srs<'c> {
    pub fn new(ctx: &'c Context, powers_of_g: &[G1Affine], powers_of_gamma_g: &[(usize, G1Affine)]) -> Result<Self, Error> {
        let mut g = Vec::with_capacity(12 * powers_of_g.len());
        powers_of_g.iter().for_each(|p| g1_limbs(p, &mut g));
        let mut gam = Vec::with_capacity(12 * powers_of_gamma_g.len());
        let mut idx = Vec::with_capacity(powers_of_gamma_g.len());
        for (i, p) in powers_of_gamma_g {
            idx.push(*i as u64);
            g1_limbs(p, &mut gam);
        }
        let mut raw = std::ptr::null_mut();
        check(unsafe {
            ffi::b2m_srs_create(ctx.raw, ffi::B2M_CURVE_BLS12_381, g.as_ptr(), powers_of_g.len(), gam.as_ptr(), idx.as_ptr(), idx.len(), 0,
                                &mut raw)
        })?;
        Ok(Srs { raw, _ctx: ctx })
    }
    /// sum_i scalars[i] * powers_of_g[offset + i]: the body of `KZG10::commit`'s MSM (`offset` = `max_degree - bound` for
    /// shifted powers).
    pub fn msm(&self, offset: usize, scalars: &[<Fr as PrimeField>::BigInt]) -> Result<G1Projective, Error> {
        let flat: Vec<u64> = scalars.iter().flat_map(|s| s.0.iter().copied().collect::<Vec<_>>()).collect();
        let mut out = [0u64; 12];
        let mut inf: c_int = 0;
        check(unsafe { ffi::b2m_srs_msm(self.raw, offset, flat.as_ptr(), scalars.len(), out.as_mut_ptr(), &mut inf) })?;
        Ok(g1_from_limbs(&out, inf != 0).into_projective())
    }
}
impl<'c> Drop for Srs<'c> {
    fn drop(&mut self) {
        unsafe { ffi::b2m_srs_destroy(self.raw) }
    }
}

/// CSR image of `Matrix<F> = Vec<Vec<(F, usize)>>` (`src/ahp/indexer.rs:81`).
struct Csr {
    row_ptr: Vec<u64>,
    col: Vec<u64>,
    coeff: Vec<u64>,
}
fn to_csr(m: &[Vec<(Fr, usize)>]) -> Csr {
    let mut row_ptr = vec![0u64];
    let (mut col, mut coeff) = (Vec::new(), Vec::new());
    for row in m {
        for (v, c) in row {
            col.push(*c as u64);
            coeff.extend_from_slice(&(v.0).0);
        }
        row_ptr.push(col.len() as u64);
    }
    Csr { row_ptr, col, coeff }
}

/// `pad_input_for_indexer_and_prover` (`src/ahp/constraint_systems.rs:45-58`): public inputs up to a power of two with zeros.
fn pad_input_for_indexer_and_prover(cs: &ark_relations::r1cs::ConstraintSystemRef<Fr>) -> Result<(), Error> {
    let n_in = cs.num_instance_variables();
    for _ in n_in..n_in.next_power_of_two() {
        cs.new_input_variable(|| Ok(Fr::zero())).map_err(Error::Synthesis)?;
    }
    Ok(())
}
/// `make_matrices_square_for_indexer` / `_for_prover` (`src/ahp/constraint_systems.rs:60-81`), applied AFTER `finalize()`:
/// dummy constraints or dummy witnesses (value one) until #constraints == #variables.
fn make_matrices_square(cs: &ark_relations::r1cs::ConstraintSystemRef<Fr>) -> Result<(), Error> {
    let vars = cs.num_instance_variables() + cs.num_witness_variables();
    let cons = cs.num_constraints();
    if vars > cons {
        for _ in cons..vars {
            cs.enforce_constraint(lc!(), lc!(), lc!()).map_err(Error::Synthesis)?;
        }
    } else {
        for _ in vars..cons {
            cs.new_witness_variable(|| Ok(Fr::one())).map_err(Error::Synthesis)?;
        }
    }
    Ok(())
}

/// `Marlin::index` on the device: index polynomials, their commitments and the verifier-key bytes.
pub struct IndexProverKey<'s> {
    raw: *mut ffi::b2m_index,
    /// `ToBytes` image of `index_vk` (index_info || index_comms): what the transcript absorbs and what
    /// `IndexVerifierKey` is rebuilt from on the Rust side.
    pub vk_bytes: Vec<u8>,
    _srs: &'s Srs<'s>,
}
impl<'s> IndexProverKey<'s> {
    pub fn new<C: ConstraintSynthesizer<Fr>>(srs: &'s Srs<'s>, circuit: C, sonic: bool) -> Result<Self, Error> {
        let cs = ConstraintSystem::<Fr>::new_ref();
        cs.set_optimization_goal(OptimizationGoal::Weight);
        cs.set_mode(SynthesisMode::Setup);
        circuit.generate_constraints(cs.clone()).map_err(Error::Synthesis)?;
        // the reference's order (src/ahp/indexer.rs:160-166): pad the public input, finalize (which may OUTLINE linear
        // combinations into new witnesses and constraints), and only then square the matrices
        pad_input_for_indexer_and_prover(&cs)?;
        cs.finalize();
        make_matrices_square(&cs)?;
        let m: ConstraintMatrices<Fr> = cs.to_matrices().expect("matrices in setup mode");
        let (a, b, c) = (to_csr(&m.a), to_csr(&m.b), to_csr(&m.c));
        let view = |x: &Csr| ffi::b2m_matrix { row_ptr: x.row_ptr.as_ptr(), col: x.col.as_ptr(), coeff: x.coeff.as_ptr() };
        let (ma, mb, mc) = (view(&a), view(&b), view(&c));
        let mut raw = std::ptr::null_mut();
        let pc = if sonic { ffi::B2M_PC_SONIC_KZG10 } else { ffi::B2M_PC_MARLIN_KZG10 };
        check(unsafe {
            ffi::b2m_index_create(srs.raw, pc, m.num_constraints, m.num_instance_variables + m.num_witness_variables,
                                  m.num_instance_variables, &ma, &mb, &mc, &mut raw)
        })?;
        let mut buf = vec![0u8; 4096];
        let mut len = 0usize;
        check(unsafe { ffi::b2m_index_vk_bytes(raw, buf.as_mut_ptr(), buf.len(), &mut len) })?;
        buf.truncate(len);
        Ok(IndexProverKey { raw, vk_bytes: buf, _srs: srs })
    }

    /// `Marlin::prove`: synthesises the witness on the CPU (the circuit is the caller's code), proves on the GPU and
    /// returns the reference's own `Proof` type.  `zk_rng` advances exactly as it would in the reference.
    pub fn prove<C, PC>(&self, circuit: C, zk_rng: &mut ChaCha12Rng) -> Result<ark_marlin::Proof<Fr, PC>, Error>
    where
        C: ConstraintSynthesizer<Fr>,
        PC: ark_poly_commit::PolynomialCommitment<Fr, ark_poly::univariate::DensePolynomial<Fr>>,
    {
        // fast path: `ark_std::test_rng()` / StdRng streams are described by (key, word position) and sampled on the device
        let mut rng = ffi::b2m_rng { kind: 12, key: zk_rng.get_seed(), word_pos: zk_rng.get_word_pos() as u64, next_u64: None,
                                     state: std::ptr::null_mut() };
        let proof = self.prove_with(circuit, &mut rng)?;
        zk_rng.set_word_pos(rng.word_pos as u128);
        Ok(proof)
    }

    /// `Marlin::prove` with ANY `RngCore` as `zk_rng` (reference src/lib.rs:154): every draw goes through the callback form of
    /// `b2m_rng`, in the reference's order.
    pub fn prove_with_rng<C, PC, R: rand_core::RngCore>(&self, circuit: C, zk_rng: &mut R) -> Result<ark_marlin::Proof<Fr, PC>, Error>
    where
        C: ConstraintSynthesizer<Fr>,
        PC: ark_poly_commit::PolynomialCommitment<Fr, ark_poly::univariate::DensePolynomial<Fr>>,
    {
        let mut dynrng: &mut dyn rand_core::RngCore = zk_rng;
        let mut rng = pc::callback_rng(&mut dynrng);
        self.prove_with(circuit, &mut rng)
    }

    fn prove_with<C, PC>(&self, circuit: C, rng: &mut ffi::b2m_rng) -> Result<ark_marlin::Proof<Fr, PC>, Error>
    where
        C: ConstraintSynthesizer<Fr>,
        PC: ark_poly_commit::PolynomialCommitment<Fr, ark_poly::univariate::DensePolynomial<Fr>>,
    {
        let cs = ConstraintSystem::<Fr>::new_ref();
        cs.set_optimization_goal(OptimizationGoal::Weight);
        // construct_matrices: true like the reference (src/ahp/prover.rs:220-222): finalize() outlines linear combinations
        // only when it builds the matrices, and the outlining adds witnesses that must exist in the assignment
        cs.set_mode(SynthesisMode::Prove { construct_matrices: true });
        circuit.generate_constraints(cs.clone()).map_err(Error::Synthesis)?;
        pad_input_for_indexer_and_prover(&cs)?;  // src/ahp/prover.rs:227-229: same order as the indexer
        cs.finalize();
        make_matrices_square(&cs)?;
        let inner = cs.borrow().expect("constraint system");
        let x = fr_mont_limbs(&inner.instance_assignment);
        let w = fr_mont_limbs(&inner.witness_assignment);
        let mut buf = vec![0u8; 4096];
        let mut len = 0usize;
        check(unsafe {
            ffi::b2m_prove(self.raw, x.as_ptr(), inner.instance_assignment.len(), w.as_ptr(), inner.witness_assignment.len(), rng,
                           buf.as_mut_ptr(), buf.len(), &mut len)
        })?;
        ark_marlin::Proof::<Fr, PC>::deserialize(&buf[..len]).map_err(Error::Serialization)
    }
}
impl<'s> Drop for IndexProverKey<'s> {
    fn drop(&mut self) {
        unsafe { ffi::b2m_index_destroy(self.raw) }
    }
}

/// The pairing engine the proofs are for (re-exported so callers name one type for both sides).
pub type Engine = Bls12_381;
