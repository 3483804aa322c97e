//! bzhalo2-sys -- raw bindings of include/bzhalo2.h plus the safe wrappers a patched `halo2_proofs 0.2.0` calls.
//! UNBUILT SOURCE (no cargo/rustc in this environment).  Layout contract: `pasta_curves::{Fp, Fq}` are
//! `#[repr(transparent)]` over `[u64; 4]` in Montgomery form, which is exactly the 32-byte element of the ABI, so
//! `&[Fp]` is passed as `*const c_void` with no conversion.  Affine points are marshalled explicitly as x || y
//! (the Rust struct layout of `EpAffine` is not `repr(C)`), identity = 64 zero bytes.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct bz_ctx { _p: [u8; 0] }
#[repr(C)] pub struct bz_params { _p: [u8; 0] }
#[repr(C)] pub struct bz_pk { _p: [u8; 0] }

#[repr(C)]
pub struct bz_token { pub op: u32, pub a: u32, pub b: i32 }

#[repr(C)]
pub struct bz_circuit {
    pub k: u32, pub num_advice: u32, pub num_fixed: u32, pub num_instance: u32, pub degree: u32, pub blinding_factors: u32,
    pub n_advice_queries: u32, pub advice_queries: *const i32,
    pub n_fixed_queries: u32, pub fixed_queries: *const i32,
    pub n_instance_queries: u32, pub instance_queries: *const i32,
    pub n_perm_columns: u32, pub perm_columns: *const u32,
    pub n_constants: u32, pub constants: *const c_void,
    pub n_tokens: u32, pub tokens: *const bz_token,
    pub n_gate_polys: u32, pub gate_poly_offsets: *const u32,
    pub n_lookups: u32, pub lookup_input_counts: *const u32, pub lookup_table_counts: *const u32,
    pub lookup_expr_offsets: *const u32,
    pub vk_transcript_repr: [u8; 32],
}

extern "C" {
    pub fn bz_ctx_create(device: c_int, stream: *mut c_void, out: *mut *mut bz_ctx) -> c_int;
    pub fn bz_ctx_destroy(ctx: *mut bz_ctx);
    pub fn bz_last_error(ctx: *mut bz_ctx) -> *const c_char;
    // arithmetic.rs
    pub fn bz_best_multiexp(ctx: *mut bz_ctx, curve: c_int, coeffs: *const c_void, bases: *const c_void, n: u64, out_jac: *mut c_void) -> c_int;
    pub fn bz_best_fft(ctx: *mut bz_ctx, field: c_int, a: *mut c_void, omega: *const c_void, log_n: u32) -> c_int;
    // poly/domain.rs
    pub fn bz_lagrange_to_coeff(ctx: *mut bz_ctx, field: c_int, a: *mut c_void, k: u32) -> c_int;
    pub fn bz_coeff_to_extended(ctx: *mut bz_ctx, field: c_int, coeffs: *const c_void, out: *mut c_void, k: u32, extended_k: u32) -> c_int;
    pub fn bz_extended_to_coeff(ctx: *mut bz_ctx, field: c_int, a: *mut c_void, extended_k: u32) -> c_int;
    // poly/commitment.rs
    pub fn bz_params_create(ctx: *mut bz_ctx, k: u32, curve: c_int, g: *const c_void, g_lagrange: *const c_void, w: *const c_void, u: *const c_void, window_bits: c_int, out: *mut *mut bz_params) -> c_int;
    pub fn bz_params_destroy(p: *mut bz_params);
    pub fn bz_params_commit(ctx: *mut bz_ctx, p: *mut bz_params, lagrange_basis: c_int, poly: *const c_void, blind: *const c_void, out_affine: *mut c_void) -> c_int;
    pub fn bz_params_commit_batch_dev(ctx: *mut bz_ctx, p: *mut bz_params, lagrange_basis: c_int, d_polys: *const c_void, d_blinds: *const c_void, count: u32, d_out_affine: *mut c_void) -> c_int;
    // plonk/prover.rs
    pub fn bz_pk_create(ctx: *mut bz_ctx, p: *mut bz_params, cs: *const bz_circuit, fixed_values: *const c_void, sigma_values: *const c_void, out: *mut *mut bz_pk) -> c_int;
    pub fn bz_pk_create_from_assembly(ctx: *mut bz_ctx, p: *mut bz_params, cs: *const bz_circuit, fixed_values: *const c_void, mapping: *const u32, out: *mut *mut bz_pk) -> c_int;
    pub fn bz_pk_vk_commitments(ctx: *mut bz_ctx, pk: *mut bz_pk, fixed_commitments: *mut c_void, perm_commitments: *mut c_void) -> c_int;
    pub fn bz_pk_destroy(pk: *mut bz_pk);
    pub fn bz_pk_num_random(pk: *const bz_pk) -> u32;
    pub fn bz_pk_proof_size(pk: *const bz_pk) -> u32;
    pub fn bz_pk_quotient_muls(pk: *const bz_pk, tier: u32, points: *mut u32) -> u32;
    // poly.rs `batch_invert_assigned`: Assigned<F> = numerator / denominator per cell (Zero = (0, 1), Trivial(x) = (x, 1))
    pub fn bz_batch_invert_assigned(ctx: *mut bz_ctx, field: c_int, numerators: *const c_void, denominators: *const c_void, out: *mut c_void, n: u64) -> c_int;
    pub fn bz_batch_invert_assigned_dev(ctx: *mut bz_ctx, field: c_int, d_numerators: *const c_void, d_denominators: *const c_void, d_out: *mut c_void, n: u64) -> c_int;
    // poly/commitment.rs `Params::new`, pasta_curves `hash_to_curve`
    pub fn bz_params_new(ctx: *mut bz_ctx, k: u32, curve: c_int, g: *mut c_void, g_lagrange: *mut c_void, w: *mut c_void, u: *mut c_void) -> c_int;
    pub fn bz_hash_to_curve(ctx: *mut bz_ctx, curve: c_int, domain_prefix: *const c_char, messages: *const c_void, msg_len: u32, count: u64, out_affine: *mut c_void) -> c_int;
    pub fn bz_points_compress(ctx: *mut bz_ctx, curve: c_int, affine: *const c_void, n: u64, out32: *mut c_void) -> c_int;
    pub fn bz_points_decompress(ctx: *mut bz_ctx, curve: c_int, in32: *const c_void, n: u64, out_affine: *mut c_void, status: *mut u8) -> c_int;
    // plonk/verifier.rs
    pub fn bz_verify_proofs(ctx: *mut bz_ctx, pk: *mut bz_pk, batch: u32, instances: *const c_void, instance_lens: *const u32, instance_stride: u32, proofs: *const c_void, proof_len: u32, results: *mut u8) -> c_int;
    pub fn bz_create_proofs(ctx: *mut bz_ctx, pk: *mut bz_pk, batch: u32, instances: *const c_void, instance_lens: *const u32, instance_stride: u32, advice: *const c_void, rand_wide: *const c_void, proofs: *mut c_void) -> c_int;
}

use pasta_curves::{arithmetic::CurveAffine, vesta, Fp};
use ff::Field;

/// `arithmetic::best_multiexp::<vesta::Affine>` -- what the patched arithmetic.rs forwards to.
pub unsafe fn best_multiexp_vesta(ctx: *mut bz_ctx, coeffs: &[Fp], bases: &[vesta::Affine]) -> Result<[u8; 96], c_int> {
    assert_eq!(coeffs.len(), bases.len());
    let mut flat = vec![0u8; bases.len() * 64];          // x || y in memory (Montgomery) form, identity = zeros
    for (i, b) in bases.iter().enumerate() {
        if let Some(c) = Option::<pasta_curves::arithmetic::Coordinates<vesta::Affine>>::from(b.coordinates()) {
            flat[i * 64..i * 64 + 32].copy_from_slice(std::slice::from_raw_parts(c.x() as *const _ as *const u8, 32));
            flat[i * 64 + 32..i * 64 + 64].copy_from_slice(std::slice::from_raw_parts(c.y() as *const _ as *const u8, 32));
        }
    }
    let mut out = [0u8; 96];
    let rc = bz_best_multiexp(ctx, 0, coeffs.as_ptr() as *const c_void, flat.as_ptr() as *const c_void, coeffs.len() as u64, out.as_mut_ptr() as *mut c_void);
    if rc == 0 { Ok(out) } else { Err(rc) }
}

/// Pre-draw the RNG words `create_proof` would consume, in protocol order (the count is shape-only:
/// `bz_pk_num_random`).  `Fp::random(rng)` = `from_u512` of 8 x `next_u64()`, little-endian limbs, so handing the raw
/// 64 bytes to the device (which performs the same reduction) reproduces the reference's field elements bit for bit.
pub fn predraw<R: rand_core_shim::RngCore>(rng: &mut R, n: usize) -> Vec<u8> {
    let mut out = vec![0u8; n * 64];
    for chunk in out.chunks_mut(8) { chunk.copy_from_slice(&rng.next_u64().to_le_bytes()); }
    out
}
pub mod rand_core_shim { pub trait RngCore { fn next_u64(&mut self) -> u64; } }
