// Tail of halo2_proofs/src/plonk/prover.rs::create_proof (tag v2023_04_20) after witness synthesis, for
// Scheme = KZGCommitmentScheme<Bn256>, P = ProverGWC, T = EvmTranscript (what zero_g instantiates, src/wnn.rs:242-259).
//
//   advice:   Vec<Polynomial<Fr, LagrangeCoeff>>  -- WitnessCollection after batch_invert_assigned, one per advice column
//   instance: &[&[Fr]]                            -- the instance columns of the single circuit
//
// The backend blinds the last `blinding_factors + 1` rows itself with the caller's RNG, in upstream's draw order, runs
// every round, and returns the bytes upstream would have written to `transcript`.
use crate::arithmetic::ZG;
use zg_b200_sys::{rng_fill, zg_create_proof, ZG_E_SYNTH};

pub(crate) fn prove_on_device<R: rand_core::RngCore>(
    pk_handle: *mut zg_b200_sys::zg_pk,
    advice: &[Vec<halo2curves::bn256::Fr>],
    instance: &[&[halo2curves::bn256::Fr]],
    mut rng: R,
) -> Result<Vec<u8>, crate::plonk::Error> {
    let advice_ptrs: Vec<_> = advice.iter().map(|c| c.as_ptr()).collect();
    let inst_ptrs: Vec<_> = instance.iter().map(|c| c.as_ptr()).collect();
    let inst_lens: Vec<usize> = instance.iter().map(|c| c.len()).collect();
    let mut proof = vec![0u8; 1 << 16];
    let mut len = 0usize;
    let rc = ZG.with(|c| unsafe {
        zg_create_proof(
            c.raw(), pk_handle, advice_ptrs.as_ptr(), inst_ptrs.as_ptr(), inst_lens.as_ptr(),
            rng_fill::<R>, &mut rng as *mut R as *mut std::os::raw::c_void,
            proof.as_mut_ptr(), proof.len(), &mut len,
        )
    });
    match rc {
        0 => { proof.truncate(len); Ok(proof) }
        ZG_E_SYNTH => Err(crate::plonk::Error::ConstraintSystemFailure),   // lookup input not in table, identity commitment
        _ => Err(crate::plonk::Error::Synthesis),
    }
}
// In create_proof: `transcript.write_bytes(&prove_on_device(..)?)`-equivalent: EvmTranscript's writer is a Vec<u8>, so the
// fork appends the returned bytes to the transcript's stream; `transcript.finalize()` (src/wnn.rs:260) then returns them.
