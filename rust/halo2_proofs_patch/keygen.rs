// Additions to halo2_proofs/src/plonk/keygen.rs (tag v2023_04_20) for the whole-proof integration level: after the
// upstream body of `keygen_pk` has produced `fixed_values` (Vec<Polynomial<Fr, LagrangeCoeff>>, selectors already
// compressed by `keygen_vk`) and the permutation `Assembly`, the key is ALSO made resident on the GPU.  The upstream
// ProvingKey keeps every field it has (write_keys / read_pk of zero_g, src/io.rs:159-170, keep working); the fork adds
// one field, `pub(crate) device: Option<zg_b200_sys::ProvingKeyHandle>`.
//
//   cs      : &ConstraintSystem<Fr>   -- after compress_selectors
//   mapping : the permutation Assembly's `mapping: Vec<Vec<(usize, usize)>>`, one Vec per permutation column
//
// `serialize_cs` is the Rust twin of 0g-halo2_b200/zg_b200/plonk/serialize.py: the u32 word stream documented there
// (magic, column counts, degree, blinding factors, the three query tables, permutation columns, RPN programs of every
// gate polynomial and lookup expression) plus the constant pool.  Expression -> RPN is a post-order walk:
// Constant -> OP_CONST, Fixed/Advice/Instance(query) -> OP_*, Negated -> OP_NEG, Sum(a, Negated(b)) -> OP_SUB,
// Sum -> OP_ADD, Product -> OP_MUL, Scaled(a, f) -> OP_SCALE.
use crate::arithmetic::ZG;
use halo2curves::bn256::{Fr, G1Affine};
use zg_b200_sys::{zg_pk, zg_pk_desc, zg_pk_load, zg_srs_load};

pub(crate) fn load_key_on_device(
    k: u32,
    g: &[G1Affine],            // ParamsKZG::g
    g_lagrange: &[G1Affine],   // ParamsKZG::g_lagrange
    cs_words: &[u32],
    constants: &[Fr],
    fixed_values: &[&[Fr]],
    mapping: &[Vec<(usize, usize)>],
    transcript_repr: Fr,       // vk.transcript_repr()
) -> Result<*mut zg_pk, crate::plonk::Error> {
    let n = 1usize << k;
    let mut flat = Vec::<u32>::with_capacity(mapping.len() * n * 2);
    for col in mapping {
        for &(c, r) in col {
            flat.push(c as u32);
            flat.push(r as u32);
        }
    }
    let fixed_ptrs: Vec<*const Fr> = fixed_values.iter().map(|c| c.as_ptr()).collect();
    let desc = zg_pk_desc {
        k,
        cs_words: cs_words.as_ptr(),
        cs_nwords: cs_words.len(),
        constants: constants.as_ptr(),
        n_constants: constants.len(),
        fixed: fixed_ptrs.as_ptr(),
        perm_mapping: flat.as_ptr(),
        transcript_repr,
        sigma_values: std::ptr::null(),
    };
    let mut handle: *mut zg_pk = std::ptr::null_mut();
    let rc = ZG.with(|c| unsafe {
        let rc = zg_srs_load(c.raw(), k, g.as_ptr(), g_lagrange.as_ptr());
        if rc != 0 { return rc; }
        zg_pk_load(c.raw(), &desc, &mut handle)
    });
    if rc != 0 { return Err(crate::plonk::Error::Synthesis); }
    Ok(handle)
}
// The commitments the device derives (zg_pk_commitments) equal the ones upstream's keygen_vk computed on the CPU; a debug
// build of the fork asserts it.
