//! Seeded dump of the UNMODIFIED reference prover, to pin the oracle ("parity unpinned").
//! Build inside the zero_g workspace:  cargo run --release --example ref_dump -- models/<model>.hdf5 benches/example_image_7.png <k> out.json
//! Uses the SRS secret and XorShift seed of tests/test_gpu_prover.py, so `out.json`'s proof bytes must equal the
//! bytes of oracle/halo2_ref.py and of zg_create_proof (modulo vk.transcript_repr, which is dumped as well and is
//! an input of both).
use halo2_proofs::halo2curves::bn256::{Bn256, Fr};
use halo2_proofs::halo2curves::ff::{Field, PrimeField};
use halo2_proofs::poly::kzg::commitment::ParamsKZG;
use rand_core::SeedableRng;
use rand_xorshift::XorShiftRng;                 // add rand_xorshift = "0.3" to [dev-dependencies]
use std::path::Path;
use zero_g::{io::{load_grayscale_image, load_wnn}};

fn main() {
    let a: Vec<String> = std::env::args().collect();
    let (wnn, img, k) = (load_wnn(Path::new(&a[1])).unwrap(), load_grayscale_image(Path::new(&a[2])).unwrap(), a[3].parse::<u32>().unwrap());
    // test SRS with the known trapdoor of the parity tests: ParamsKZG::setup takes an RNG, so the secret is injected by
    // an RNG whose first Fr::random is s.  (unsafe_setup_with_s exists on later tags; on v2023_04_20 patch setup.)
    let s = Fr::from_u128(0x1F3C5A7B9D2E4F60718293A4B5C6D7E8u128);
    let params = ParamsKZG::<Bn256>::unsafe_setup_with_s(k, s);
    let pk = wnn.generate_proving_key(&params);
    // Wnn::proof uses OsRng (src/wnn.rs:250); the dumper calls create_proof with the seeded RNG instead
    let rng = XorShiftRng::from_seed(core::array::from_fn(|i| i as u8));
    let (proof, outputs) = zero_g::wnn::proof_with_rng(&wnn, &pk, &params, &img, rng);
    let hex = |b: &[u8]| b.iter().map(|x| format!("{x:02x}")).collect::<String>();
    let vk = pk.get_vk();
    println!("{{\"k\":{k},\"transcript_repr\":\"{}\",\"fixed_commitments\":[{}],\"outputs\":[{}],\"proof\":\"{}\"}}",
        hex(vk.transcript_repr().to_repr().as_ref()),
        vk.fixed_commitments().iter().map(|c| format!("\"{:?}\"", c)).collect::<Vec<_>>().join(","),
        outputs.iter().map(|o| format!("\"{}\"", hex(o.to_repr().as_ref()))).collect::<Vec<_>>().join(","),
        hex(&proof));
    let _ = Fr::ONE;
}
