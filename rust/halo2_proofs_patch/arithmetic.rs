// Replacement bodies for halo2_proofs/src/arithmetic.rs (tag v2023_04_20).  Signatures are upstream's.
use zg_b200_sys::{Context, ZG_BASIS_LAGRANGE, ZG_BASIS_MONOMIAL};

thread_local! {
    pub(crate) static ZG: Context = Context::new(
        std::env::var("ZG_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0)
    ).expect("zg_b200: no CUDA device (this backend has no CPU fallback)");
}

/// `best_multiexp` is only ever called with `bases` = a prefix of `ParamsKZG::g` or `::g_lagrange`
/// (commit / commit_lagrange / the GWC witnesses); `ParamsKZG` records which basis a slice belongs to and the
/// commitment methods call this instead of the generic function.
pub(crate) fn multiexp_on_basis(coeffs: &[halo2curves::bn256::Fr], lagrange: bool) -> halo2curves::bn256::G1 {
    ZG.with(|c| c.msm(if lagrange { ZG_BASIS_LAGRANGE } else { ZG_BASIS_MONOMIAL }, coeffs))
        .unwrap_or_else(|e| panic!("zg_msm: {}", e.message))
}

/// `best_fft::<Fr, Fr>`: in place, natural order in and out.
pub fn best_fft(a: &mut [halo2curves::bn256::Fr], omega: halo2curves::bn256::Fr, log_n: u32) {
    ZG.with(|c| c.fft(a, omega, log_n)).unwrap_or_else(|e| panic!("zg_ntt: {}", e.message))
}
