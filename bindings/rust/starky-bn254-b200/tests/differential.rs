//! THE parity pin: proofs made by the B200 prover / its byte-identical CPU oracle are given
//! to the UNMODIFIED `starky::verifier::verify_stark_proof` of starky 0.1.1 @ InternetMaximalism/plonky2 541e127.
//!
//!   1. `python tools/export_proofs_for_rust.py`  (CPU only: gcc + the oracle; on a GPU box add --gpu to export GPU proofs instead)
//!      writes  tests/data/fq_128_{B,A}_{pad,hack}.proof  -- FqExpStark, num_io = 128, seed 0x5EED0000, one file per setting of
//!      U1 (generator pair B = (7, 1753635133440165772) / A = (14293326489335486720, 7277203076849721926)) x U3 (padded / degree hack);
//!   2. `SBN_LIB_DIR=$PWD/starky-bn254_b200 cargo test -p starky-bn254-b200 --release -- --nocapture`.
//! Exactly one of the four files must verify: that names the conventions of this starky version (U1, U3) and, with them, pins
//! U2 (any PoW witness is accepted), U5 / U6 (challenge order, sponge mode: a wrong one fails every file).  `gpu_end_to_end` then
//! checks the live path on a machine with a GPU.
use plonky2::plonk::config::{GenericConfig, PoseidonGoldilocksConfig};
use starky::verifier::verify_stark_proof;
use starky_bn254::fields::fq::exp::FqExpStark;
use starky_bn254_b200::to_stark_proof;
use starky_bn254_b200_sys::decode_proof;

const D: usize = 2;
type C = PoseidonGoldilocksConfig;
type F = <C as GenericConfig<D>>::F;

#[test]
fn exported_proofs_against_the_unmodified_verifier() {
    let dir = std::path::Path::new(env!("CARGO_MANIFEST_DIR")).join("tests/data");
    let stark = FqExpStark::<F, D>::new(128);
    let config = stark.config();
    let mut accepted = vec![];
    for name in ["fq_128_B_pad", "fq_128_A_pad", "fq_128_B_hack", "fq_128_A_hack"] {
        let path = dir.join(format!("{name}.proof"));
        let Ok(bytes) = std::fs::read(&path) else { eprintln!("{} missing: run tools/export_proofs_for_rust.py first", path.display()); continue };
        let proof = to_stark_proof::<F, C, D>(&decode_proof(&bytes)).expect("wire format");
        assert_eq!(proof.proof.recover_degree_bits(&config), 16);
        match verify_stark_proof(stark, proof, &config) {
            Ok(()) => { println!("{name}: ACCEPTED by starky::verifier::verify_stark_proof"); accepted.push(name); }
            Err(e) => println!("{name}: rejected ({e})"),
        }
    }
    assert_eq!(accepted.len(), 1, "exactly one (U1, U3) setting must verify; accepted: {accepted:?}");
}

/// On a machine with a GPU: random inputs as in the reference's own test (src/fields/fq/exp.rs), proved on the GPU, verified by starky.
#[test]
#[ignore = "needs a CUDA device and libstarkybn254_b200.so"]
fn gpu_end_to_end() {
    use ark_bn254::Fq;
    use ark_ff::Field as _;
    use ark_std::UniformRand;
    use starky_bn254::fields::fq::exp::FqExpIONative;
    let mut rng = rand::thread_rng();
    let inputs: Vec<FqExpIONative> = (0..128).map(|_| {
        let exp_val: [u32; 8] = rand::random();
        let (x, offset) = (Fq::rand(&mut rng), Fq::rand(&mut rng));
        let e: Vec<u64> = exp_val.chunks(2).map(|c| c[0] as u64 | (c[1] as u64) << 32).collect();
        FqExpIONative { x, offset, exp_val, output: offset * x.pow(e) }
    }).collect();
    let stark = FqExpStark::<F, D>::new(128);
    let config = stark.config();
    let proof = starky_bn254_b200::generate_trace_and_prove_gpu::<F, C, _, D>(stark, &config, &inputs).unwrap();
    verify_stark_proof(stark, proof, &config).unwrap();
}
