//! Seeded dump of the UNMODIFIED reference prover: pins oracle/halo2_ref.py and the GPU prover against the real crates
//! ("parity unpinned").  No patch to zero_g or halo2_proofs is needed: `Wnn::get_circuit`,
//! `Wnn::generate_proving_key` and `Wnn::predict` are public (/root/reference/src/wnn.rs:152, :187, :222) and
//! `create_proof` is generic over its RNG (src/wnn.rs:242-259 passes OsRng; this tool passes a seeded XorShiftRng).
//!
//! Put this file at zero_g/examples/ref_dump.rs, add to zero_g/Cargo.toml
//!     [dev-dependencies]
//!     rand_xorshift = "0.3"
//! and run
//!     cargo run --release --example ref_dump -- models/model_28input_256entry_1hash_1bpi.hdf5 benches/example_image_7.png 14 > ref_dump_tiny.json
//! then `python scripts/compare_ref_dump.py ref_dump_tiny.json` in this repository.
//!
//! The SRS: ParamsKZG::setup(k, rng) draws s = Fr::random(rng) first (halo2_proofs v2023_04_20, poly/kzg/commitment.rs);
//! `Fr::random` is from_u512 of eight next_u64 draws (halo2curves 0.3.3), so an RNG that yields the four limbs of the
//! test secret followed by zeros makes s the secret of tests/test_gpu_prover.py without any `unsafe_setup` API.
use halo2_proofs::halo2curves::bn256::{Bn256, Fr, G1Affine};
use halo2_proofs::halo2curves::ff::PrimeField;
use halo2_proofs::plonk::create_proof;
use halo2_proofs::poly::kzg::commitment::{KZGCommitmentScheme, ParamsKZG};
use halo2_proofs::poly::kzg::multiopen::ProverGWC;
use halo2_proofs::transcript::TranscriptWriterBuffer;
use rand_core::{RngCore, SeedableRng};
use rand_xorshift::XorShiftRng;
use snark_verifier::system::halo2::transcript::evm::EvmTranscript;
use std::path::Path;
use zero_g::{load_grayscale_image, load_wnn};

/// yields the given u64 words, then zeros
struct FixedRng { words: Vec<u64>, pos: usize }
impl RngCore for FixedRng {
    fn next_u32(&mut self) -> u32 { self.next_u64() as u32 }
    fn next_u64(&mut self) -> u64 {
        let w = self.words.get(self.pos).copied().unwrap_or(0);
        self.pos += 1;
        w
    }
    fn fill_bytes(&mut self, dest: &mut [u8]) {
        for chunk in dest.chunks_mut(8) {
            let w = self.next_u64().to_le_bytes();
            chunk.copy_from_slice(&w[..chunk.len()]);
        }
    }
    fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), rand_core::Error> { self.fill_bytes(dest); Ok(()) }
}

fn hex(b: &[u8]) -> String { b.iter().map(|x| format!("{x:02x}")).collect() }
fn point(p: &G1Affine) -> String { format!("[\"{}\",\"{}\"]", hex(p.x.to_repr().as_ref()), hex(p.y.to_repr().as_ref())) }

fn main() {
    let a: Vec<String> = std::env::args().collect();
    let wnn = load_wnn(Path::new(&a[1])).unwrap();
    let img = load_grayscale_image(Path::new(&a[2])).unwrap();
    let k: u32 = a[3].parse().unwrap();
    // SRS_SECRET of bench.py / the parity tests, little-endian limbs
    let secret: u128 = 0x1F3C5A7B9D2E4F60718293A4B5C6D7E8;
    let params = ParamsKZG::<Bn256>::setup(k, FixedRng { words: vec![secret as u64, (secret >> 64) as u64], pos: 0 });
    let pk = wnn.generate_proving_key(&params);
    let outputs: Vec<Fr> = wnn.predict(&img).into_iter().map(Fr::from).collect();
    let circuit = wnn.get_circuit(&img);
    let rng = XorShiftRng::from_seed(core::array::from_fn(|i| i as u8));       // SEED = bytes(range(16)) in the tests
    let mut transcript = TranscriptWriterBuffer::<_, G1Affine, _>::init(Vec::new());
    create_proof::<KZGCommitmentScheme<Bn256>, ProverGWC<_>, _, _, EvmTranscript<_, _, _, _>, _>(
        &params, &pk, &[circuit], &[&[outputs.as_ref()]], rng, &mut transcript,
    ).unwrap();
    let proof = transcript.finalize();
    let vk = pk.get_vk();
    println!(
        "{{\"k\":{k},\"transcript_repr\":\"{}\",\"fixed_commitments\":[{}],\"permutation_commitments\":[{}],\"outputs\":[{}],\"proof\":\"{}\"}}",
        hex(vk.transcript_repr().to_repr().as_ref()),
        vk.fixed_commitments().iter().map(point).collect::<Vec<_>>().join(","),
        vk.permutation().commitments().iter().map(point).collect::<Vec<_>>().join(","),
        outputs.iter().map(|o| format!("\"{}\"", hex(o.to_repr().as_ref()))).collect::<Vec<_>>().join(","),
        hex(&proof)
    );
}
