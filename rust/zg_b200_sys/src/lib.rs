//! FFI surface of `libzg_b200.so` (see `include/zg_b200.h` for the contract of every entry point).
//!
//! `Fr` / `Fq` of halo2curves 0.3.3 are `#[repr(transparent)]` wrappers of `[u64; 4]` in Montgomery form, which is
//! exactly `zg_fr` / `zg_fq`; `G1Affine { x, y }` is `zg_g1_affine`, `G1 { x, y, z }` is `zg_g1`.  Slices cross the
//! boundary as pointers, with no conversion.
#![allow(non_camel_case_types)]
use halo2curves::bn256::{Fr, G1Affine, G1};
use rand_core::RngCore;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct zg_ctx { _p: [u8; 0] }
#[repr(C)] pub struct zg_pk { _p: [u8; 0] }

#[repr(C)]
pub struct zg_pk_desc {
    pub k: u32,
    pub cs_words: *const u32,
    pub cs_nwords: usize,
    pub constants: *const Fr,
    pub n_constants: usize,
    pub fixed: *const *const Fr,
    pub perm_mapping: *const u32,        // or null, with sigma_values set (a key read from a file)
    pub transcript_repr: Fr,
    pub sigma_values: *const *const Fr,
}

pub type zg_rng_fill_fn = unsafe extern "C" fn(state: *mut c_void, out: *mut u64, n: usize);

pub const ZG_BASIS_MONOMIAL: c_int = 0;
pub const ZG_BASIS_LAGRANGE: c_int = 1;
pub const ZG_E_SYNTH: c_int = -5;

extern "C" {
    pub fn zg_ctx_create(device: c_int, stream: *mut c_void, out: *mut *mut zg_ctx) -> c_int;
    pub fn zg_ctx_destroy(ctx: *mut zg_ctx);
    pub fn zg_last_error(ctx: *const zg_ctx) -> *const c_char;
    pub fn zg_sync(ctx: *mut zg_ctx) -> c_int;
    pub fn zg_srs_load(ctx: *mut zg_ctx, k: u32, g: *const G1Affine, g_lagrange: *const G1Affine) -> c_int;
    pub fn zg_msm(ctx: *mut zg_ctx, basis: c_int, scalars: *const Fr, n: usize, out: *mut G1) -> c_int;
    pub fn zg_msm_batch(ctx: *mut zg_ctx, basis: c_int, scalars: *const *const Fr, n: usize, count: usize, out: *mut G1) -> c_int;
    pub fn zg_ntt(ctx: *mut zg_ctx, a: *mut Fr, log_n: u32, omega: *const Fr) -> c_int;
    pub fn zg_lagrange_to_coeff(ctx: *mut zg_ctx, a: *mut Fr, k: u32) -> c_int;
    pub fn zg_coeff_to_extended(ctx: *mut zg_ctx, coeff: *const Fr, k: u32, ext_k: u32, out: *mut Fr) -> c_int;
    pub fn zg_extended_to_coeff(ctx: *mut zg_ctx, ext: *const Fr, k: u32, ext_k: u32, keep: usize, out: *mut Fr) -> c_int;
    pub fn zg_lookup_permute(ctx: *mut zg_ctx, a: *const Fr, s: *const Fr, usable: usize, a_perm: *mut Fr, s_perm: *mut Fr) -> c_int;
    pub fn zg_grand_product(ctx: *mut zg_ctx, num: *const Fr, den: *const Fr, len: usize, z: *mut Fr) -> c_int;
    pub fn zg_batch_invert(ctx: *mut zg_ctx, a: *mut Fr, n: usize) -> c_int;
    pub fn zg_eval_poly_batch(ctx: *mut zg_ctx, polys: *const *const Fr, n: usize, count: usize, x: *const Fr, out: *mut Fr) -> c_int;
    pub fn zg_kate_division(ctx: *mut zg_ctx, a: *const Fr, n: usize, z: *const Fr, q: *mut Fr) -> c_int;
    pub fn zg_pk_load(ctx: *mut zg_ctx, desc: *const zg_pk_desc, out: *mut *mut zg_pk) -> c_int;
    pub fn zg_pk_free(ctx: *mut zg_ctx, pk: *mut zg_pk);
    pub fn zg_pk_commitments(ctx: *mut zg_ctx, pk: *const zg_pk, fixed: *mut G1Affine, sigma: *mut G1Affine) -> c_int;
    pub fn zg_pk_set_transcript_repr(ctx: *mut zg_ctx, pk: *mut zg_pk, transcript_repr: *const Fr) -> c_int;
    pub fn zg_evaluate_h(ctx: *mut zg_ctx, pk: *mut zg_pk, advice_polys: *const *const Fr, instance_polys: *const *const Fr,
                         lookup_input_polys: *const *const Fr, lookup_table_polys: *const *const Fr,
                         lookup_product_polys: *const *const Fr, perm_product_polys: *const *const Fr,
                         challenges: *const Fr, divide: c_int, h_out: *mut Fr) -> c_int;
    pub fn zg_create_proof(ctx: *mut zg_ctx, pk: *mut zg_pk, advice: *const *const Fr, instances: *const *const Fr,
                           instance_lens: *const usize, rng: zg_rng_fill_fn, rng_state: *mut c_void,
                           proof_out: *mut u8, proof_cap: usize, proof_len: *mut usize) -> c_int;
}

/// `RngCore::next_u64` in bulk: the backend draws exactly the u64 stream upstream's `Fr::random` calls would draw
/// (8 per field element), so a seeded RNG gives the same proof as the CPU prover.
pub unsafe extern "C" fn rng_fill<R: RngCore>(state: *mut c_void, out: *mut u64, n: usize) {
    let rng = &mut *(state as *mut R);
    for w in std::slice::from_raw_parts_mut(out, n) {
        *w = rng.next_u64();
    }
}

#[derive(Debug)]
pub struct Error { pub code: c_int, pub message: String }

/// One GPU, one stream; not `Sync` (one context per host thread, as the header states).
pub struct Context { raw: *mut zg_ctx }

impl Context {
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { zg_ctx_create(device, std::ptr::null_mut(), &mut raw) };
        if rc != 0 {
            return Err(Error { code: rc, message: "zg_ctx_create failed: no CUDA device (there is no CPU fallback)".into() });
        }
        Ok(Self { raw })
    }
    pub fn raw(&self) -> *mut zg_ctx { self.raw }
    pub fn check(&self, rc: c_int) -> Result<(), Error> {
        if rc == 0 { return Ok(()); }
        let message = unsafe { CStr::from_ptr(zg_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(Error { code: rc, message })
    }
    /// `ParamsKZG::{g, g_lagrange}` -> device + fixed-base window tables (once per (context, k)).
    pub fn load_srs(&self, k: u32, g: &[G1Affine], g_lagrange: &[G1Affine]) -> Result<(), Error> {
        assert!(g.len() == 1 << k && g_lagrange.len() == 1 << k);
        self.check(unsafe { zg_srs_load(self.raw, k, g.as_ptr(), g_lagrange.as_ptr()) })
    }
    /// `arithmetic::best_multiexp(coeffs, bases)` for bases = the first `coeffs.len()` points of a loaded basis.
    pub fn msm(&self, basis: c_int, coeffs: &[Fr]) -> Result<G1, Error> {
        let mut out = G1::default();
        self.check(unsafe { zg_msm(self.raw, basis, coeffs.as_ptr(), coeffs.len(), &mut out) })?;
        Ok(out)
    }
    /// `arithmetic::best_fft(a, omega, log_n)`
    pub fn fft(&self, a: &mut [Fr], omega: Fr, log_n: u32) -> Result<(), Error> {
        assert!(a.len() == 1 << log_n);
        self.check(unsafe { zg_ntt(self.raw, a.as_mut_ptr(), log_n, &omega) })
    }
}

impl Drop for Context {
    fn drop(&mut self) { unsafe { zg_ctx_destroy(self.raw) } }
}
