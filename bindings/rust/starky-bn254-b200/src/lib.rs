//! Drop-in GPU path for the provers of qope/starky-bn254: replaces
//! ```ignore
//! let trace = stark.generate_trace(&inputs);                       // src/curves/g1/exp.rs:816
//! let pi    = stark.generate_public_inputs(&inputs);               // :817
//! let proof = prove::<F, C, _, D>(stark, &config, trace, pi.try_into().unwrap(), &mut timing)?;   // :818
//! ```
//! by `generate_trace_and_prove_gpu(stark, &inputs)?` (or `prove_gpu` for a host-made trace); the `Stark` impls and
//! `verify_stark_proof(stark, proof.clone(), &config)` (:826) stay as they are.  A generic `Stark` callback cannot cross to CUDA,
//! so every production AIR carries its identifier through `GpuStark`.
//! Written against starky 0.1.1 / plonky2 0.1.3 @ InternetMaximalism/plonky2 541e127; never compiled here (no toolchain).
use std::cell::RefCell;

use anyhow::{anyhow, Result};
use ark_bn254::{Fq, Fq12, Fq2, G1Affine, G2Affine};
use ark_ff::PrimeField as ArkPrimeField;
use plonky2::field::extension::Extendable;
use plonky2::field::goldilocks_field::GoldilocksField;
use plonky2::field::polynomial::{PolynomialCoeffs, PolynomialValues};
use plonky2::field::types::{Field, PrimeField64};
use plonky2::fri::proof::{FriInitialTreeProof, FriProof, FriQueryRound, FriQueryStep};
use plonky2::hash::hash_types::{HashOut, RichField};
use plonky2::hash::merkle_proofs::MerkleProof;
use plonky2::hash::merkle_tree::MerkleCap;
use plonky2::plonk::config::{GenericConfig, Hasher};
use plonky2::util::timing::TimingTree;
use plonky2_bn254::fields::native::MyFq12;
use starky::config::StarkConfig;
use starky::proof::{StarkOpeningSet, StarkProof, StarkProofWithPublicInputs};
use starky::stark::Stark;
use starky_bn254::curves::g1::exp::{G1ExpIONative, G1ExpStark};
use starky_bn254::curves::g2::exp::{G2ExpIONative, G2ExpStark};
use starky_bn254::fields::fq::exp::{FqExpIONative, FqExpStark};
use starky_bn254::fields::fq12::exp::{Fq12ExpIONative, Fq12ExpStark};
use starky_bn254::fields::fq12_u64::exp_u64::{Fq12ExpU64IONative, Fq12ExpU64Stark};
use starky_bn254_b200_sys as sys;

/// AIR identifier + input-record conversion of a `Stark` implementation of the reference.
pub trait GpuStark {
    const AIR: i32;
    /// `*IONative` of the reference.
    type Native;
    /// packed record of include/starky_bn254_b200.h.
    type Record: Copy;
    fn num_io(&self) -> usize;
    fn record(io: &Self::Native) -> Self::Record;
}

fn fq_words(x: &Fq) -> [u64; 4] { x.into_bigint().0 }
fn fq2_words(x: &Fq2, out: &mut [u64]) { out[..4].copy_from_slice(&fq_words(&x.c0)); out[4..8].copy_from_slice(&fq_words(&x.c1)); }
fn g2_words(p: &G2Affine) -> [u64; 16] { let mut o = [0u64; 16]; fq2_words(&p.x, &mut o[0..8]); fq2_words(&p.y, &mut o[8..16]); o }
/// the flat `MyFq12` coefficient order the reference converts to before writing columns (src/utils/utils.rs:174-183)
fn fq12_words(x: &Fq12) -> [u64; 48] {
    let m: MyFq12 = (*x).into();
    let mut o = [0u64; 48];
    for (i, c) in m.coeffs.iter().enumerate() { o[4 * i..4 * i + 4].copy_from_slice(&fq_words(c)); }
    o
}

impl<F: RichField + Extendable<D>, const D: usize> GpuStark for G1ExpStark<F, D> {
    const AIR: i32 = sys::SBN_AIR_G1_EXP;
    type Native = G1ExpIONative;
    type Record = sys::sbn_g1_exp_io;
    fn num_io(&self) -> usize { self.num_io }
    fn record(io: &G1ExpIONative) -> sys::sbn_g1_exp_io {
        let (x, o, r): (&G1Affine, &G1Affine, &G1Affine) = (&io.x, &io.offset, &io.output);
        sys::sbn_g1_exp_io { x_x: fq_words(&x.x), x_y: fq_words(&x.y), offset_x: fq_words(&o.x), offset_y: fq_words(&o.y), exp_val: io.exp_val,
                             output_x: fq_words(&r.x), output_y: fq_words(&r.y) }
    }
}
impl<F: RichField + Extendable<D>, const D: usize> GpuStark for G2ExpStark<F, D> {
    const AIR: i32 = sys::SBN_AIR_G2_EXP;
    type Native = G2ExpIONative;
    type Record = sys::sbn_g2_exp_io;
    fn num_io(&self) -> usize { self.num_io }
    fn record(io: &G2ExpIONative) -> sys::sbn_g2_exp_io {
        sys::sbn_g2_exp_io { x: g2_words(&io.x), offset: g2_words(&io.offset), exp_val: io.exp_val, output: g2_words(&io.output) }
    }
}
impl<F: RichField + Extendable<D>, const D: usize> GpuStark for FqExpStark<F, D> {
    const AIR: i32 = sys::SBN_AIR_FQ_EXP;
    type Native = FqExpIONative;
    type Record = sys::sbn_fq_exp_io;
    fn num_io(&self) -> usize { self.num_io }
    fn record(io: &FqExpIONative) -> sys::sbn_fq_exp_io {
        sys::sbn_fq_exp_io { x: fq_words(&io.x), offset: fq_words(&io.offset), exp_val: io.exp_val, output: fq_words(&io.output) }
    }
}
impl<F: RichField + Extendable<D>, const D: usize> GpuStark for Fq12ExpStark<F, D> {
    const AIR: i32 = sys::SBN_AIR_FQ12_EXP;
    type Native = Fq12ExpIONative;
    type Record = sys::sbn_fq12_exp_io;
    fn num_io(&self) -> usize { self.num_io }
    fn record(io: &Fq12ExpIONative) -> sys::sbn_fq12_exp_io {
        sys::sbn_fq12_exp_io { x: fq12_words(&io.x), offset: fq12_words(&io.offset), exp_val: io.exp_val, output: fq12_words(&io.output) }
    }
}
impl<F: RichField + Extendable<D>, const D: usize> GpuStark for Fq12ExpU64Stark<F, D> {
    const AIR: i32 = sys::SBN_AIR_FQ12_EXP_U64;
    type Native = Fq12ExpU64IONative;
    type Record = sys::sbn_fq12_exp_u64_io;
    fn num_io(&self) -> usize { self.num_io }
    fn record(io: &Fq12ExpU64IONative) -> sys::sbn_fq12_exp_u64_io {
        sys::sbn_fq12_exp_u64_io { x: fq12_words(&io.x), offset: fq12_words(&io.offset), exp_val: io.exp_val, output: fq12_words(&io.output) }
    }
}

/// `StarkConfig` -> `sbn_config`; the coset shift is the field's own (`F::coset_shift()`), so whichever generator pair the linked
/// plonky2_field has is the one the GPU uses.  `fri_degree_hack` (U3) is 0 unless SBN_FRI_DEGREE_HACK=1.
pub fn sbn_config_of<F: RichField>(config: &StarkConfig) -> sys::sbn_config {
    let fri = &config.fri_config;
    let (arity_bits, final_poly_bits) = match fri.reduction_strategy {
        plonky2::fri::reduction_strategies::FriReductionStrategy::ConstantArityBits(a, f) => (a as u32, f as u32),
        _ => panic!("only FriReductionStrategy::ConstantArityBits is supported (what standard_fast_config uses)"),
    };
    sys::sbn_config {
        security_bits: config.security_bits as u32, num_challenges: config.num_challenges as u32, rate_bits: fri.rate_bits as u32, cap_height: fri.cap_height as u32,
        pow_bits: fri.proof_of_work_bits, fri_arity_bits: arity_bits, fri_final_poly_bits: final_poly_bits, num_query_rounds: fri.num_query_rounds as u32,
        coset_shift: F::coset_shift().to_canonical_u64(),
        fri_degree_hack: std::env::var("SBN_FRI_DEGREE_HACK").map(|v| v == "1").unwrap_or(false) as u32, reserved: 0,
    }
}

thread_local! { static CTX: RefCell<Option<sys::Context>> = RefCell::new(None); }
/// one `sbn_ctx` per (thread, GPU); the device comes from SBN_DEVICE (default 0)
fn with_ctx<T>(f: impl FnOnce(&sys::Context) -> Result<T>) -> Result<T> {
    CTX.with(|c| {
        if c.borrow().is_none() {
            let dev = std::env::var("SBN_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0);
            *c.borrow_mut() = Some(sys::Context::new(dev).map_err(|e| anyhow!("{e}"))?);
        }
        f(c.borrow().as_ref().unwrap())
    })
}

/// Same shape as `starky::prover::prove`, for a trace that already exists on the host (K2-K6 on the GPU).
pub fn prove_gpu<F, C, S, const D: usize>(stark: S, config: &StarkConfig, trace_poly_values: Vec<PolynomialValues<F>>, public_inputs: Vec<F>,
                                          _timing: &mut TimingTree) -> Result<StarkProofWithPublicInputs<F, C, D>>
where F: RichField + Extendable<D>, C: GenericConfig<D, F = F>, S: Stark<F, D> + GpuStark {
    let (ncols, nrows) = (trace_poly_values.len(), trace_poly_values[0].len());
    // GoldilocksField is #[repr(transparent)] u64 but may hold non-canonical values: canonicalise on the way out
    let flat: Vec<u64> = trace_poly_values.iter().flat_map(|c| c.values.iter().map(|v| v.to_canonical_u64())).collect();
    let pis: Vec<u64> = public_inputs.iter().map(|v| v.to_canonical_u64()).collect();
    let cfg = sbn_config_of::<F>(config);
    let bytes = with_ctx(|ctx| {
        let t = ctx.upload_trace(S::AIR, stark.num_io(), &flat, ncols, nrows).map_err(|e| anyhow!("{e}"))?;
        ctx.prove(&cfg, &t, &pis).map_err(|e| anyhow!("{e}"))
    })?;
    to_stark_proof::<F, C, D>(&sys::decode_proof(&bytes))
}

/// Fast path: trace generation on the GPU too (K1-K6).  The `output` field of every record is filled in by the caller, as in the
/// reference's tests and witness generators; the library recomputes the chain and the proof's public inputs carry these outputs.
pub fn generate_trace_and_prove_gpu<F, C, S, const D: usize>(stark: S, config: &StarkConfig, inputs: &[S::Native]) -> Result<StarkProofWithPublicInputs<F, C, D>>
where F: RichField + Extendable<D>, C: GenericConfig<D, F = F>, S: Stark<F, D> + GpuStark {
    let recs: Vec<S::Record> = inputs.iter().map(S::record).collect();
    let cfg = sbn_config_of::<F>(config);
    let bytes = with_ctx(|ctx| {
        let t = ctx.generate_trace(S::AIR, &recs).map_err(|e| anyhow!("{e}"))?;
        let pis = sys::public_inputs(S::AIR, &recs).map_err(|e| anyhow!("{e}"))?;
        ctx.prove(&cfg, &t, &pis).map_err(|e| anyhow!("{e}"))
    })?;
    to_stark_proof::<F, C, D>(&sys::decode_proof(&bytes))
}

/// B independent proofs in one call (`sbn_prove_batch`): what a service proving many scalar multiplications uses.
pub fn prove_batch_gpu<F, C, S, const D: usize>(batch: &sys::Batch, config: &StarkConfig, inputs: &[&[S::Native]]) -> Result<Vec<StarkProofWithPublicInputs<F, C, D>>>
where F: RichField + Extendable<D>, C: GenericConfig<D, F = F>, S: Stark<F, D> + GpuStark {
    let recs: Vec<Vec<S::Record>> = inputs.iter().map(|b| b.iter().map(S::record).collect()).collect();
    let views: Vec<&[S::Record]> = recs.iter().map(|v| v.as_slice()).collect();
    let proofs = batch.prove(S::AIR, &sbn_config_of::<F>(config), &views, false).map_err(|e| anyhow!("{e}"))?;
    proofs.iter().map(|b| to_stark_proof::<F, C, D>(&sys::decode_proof(b))).collect()
}

/// Wire format -> the starky / plonky2 proof types, field for field.
pub fn to_stark_proof<F, C, const D: usize>(p: &sys::ProofParts) -> Result<StarkProofWithPublicInputs<F, C, D>>
where F: RichField + Extendable<D>, C: GenericConfig<D, F = F> {
    if D != 2 { return Err(anyhow!("the wire format carries quadratic-extension elements (D = 2)")); }
    let f = |v: u64| F::from_canonical_u64(v);
    let ext = |e: &sys::Ext| -> F::Extension { <F::Extension as plonky2::field::extension::FieldExtension<D>>::from_basefield_array(core::array::from_fn(|i| f(e[i]))) };
    let hash = |h: &sys::Hash| -> <C::Hasher as Hasher<F>>::Hash { hash_from_words::<F, C, D>(h) };
    let cap = |c: &Vec<sys::Hash>| MerkleCap::<F, C::Hasher>(c.iter().map(hash).collect());
    let path = |m: &sys::MerkleProof| MerkleProof::<F, C::Hasher> { siblings: m.siblings.iter().map(hash).collect() };
    let evec = |v: &Vec<sys::Ext>| -> Vec<F::Extension> { v.iter().map(ext).collect() };
    let openings = StarkOpeningSet::<F, D> {
        local_values: evec(&p.local_values), next_values: evec(&p.next_values),
        permutation_zs: p.permutation_zs.as_ref().map(evec), permutation_zs_next: p.permutation_zs_next.as_ref().map(evec),
        quotient_polys: evec(&p.quotient_polys),
    };
    let query_round_proofs = p.query_round_proofs.iter().map(|q| FriQueryRound::<F, C::Hasher, D> {
        initial_trees_proof: FriInitialTreeProof { evals_proofs: q.initial.iter().map(|(ev, mp)| (ev.iter().map(|&v| f(v)).collect(), path(mp))).collect() },
        steps: q.steps.iter().map(|s| FriQueryStep { evals: evec(&s.evals), merkle_proof: path(&s.merkle_proof) }).collect(),
    }).collect();
    let opening_proof = FriProof::<F, C::Hasher, D> {
        commit_phase_merkle_caps: p.commit_phase_merkle_caps.iter().map(cap).collect(), query_round_proofs,
        final_poly: PolynomialCoeffs::new(evec(&p.final_poly)), pow_witness: f(p.pow_witness),
    };
    let proof = StarkProof::<F, C, D> { trace_cap: cap(&p.trace_cap), permutation_zs_cap: p.permutation_zs_cap.as_ref().map(cap), quotient_polys_cap: cap(&p.quotient_polys_cap),
                                        openings, opening_proof };
    Ok(StarkProofWithPublicInputs { proof, public_inputs: p.public_inputs.iter().map(|&v| f(v)).collect::<Vec<F>>().try_into().map_err(|_| anyhow!("public input count"))? })
}
/// `HashOut<F>` from four canonical words (PoseidonGoldilocksConfig: `Hasher::Hash = HashOut<F>`).
fn hash_from_words<F: RichField + Extendable<D>, C: GenericConfig<D, F = F>, const D: usize>(h: &sys::Hash) -> <C::Hasher as Hasher<F>>::Hash {
    use plonky2::plonk::config::GenericHashOut;
    let bytes: Vec<u8> = h.iter().flat_map(|w| w.to_le_bytes()).collect();   // HashOut::from_bytes reads four LE u64 (canonical: < p)
    <<C::Hasher as Hasher<F>>::Hash as GenericHashOut<F>>::from_bytes(&bytes)
}
/// Used by the differential test only: the Goldilocks instantiation every call site of the reference makes.
pub type GoldilocksF = GoldilocksField;
pub fn _assert_hash_layout() { let _ = HashOut::<GoldilocksField>::ZERO; }
