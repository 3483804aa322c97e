"""battlezips-halo2_b200: host-side mirror of the halo2_proofs 0.2.0 prover interface over libbzhalo2.so
(hand-written sm_100a CUDA behind the C ABI of include/bzhalo2.h).

The reference's host language (Rust) is absent from this image, so the reference-facing interface is mirrored
here in Python over ctypes with the same names and argument meaning as the Rust API it stands in for
(`best_multiexp`, `best_fft`, `EvaluationDomain`, `Params`, `create_proof`, ...; call sites
/root/reference/benches/shot.rs:58-71).  There is NO CPU fallback: if the CUDA library is missing or no GPU is
visible, constructing a Context raises."""
from .binding import (Context, BzError, lib_path, load_library, build_library, EXPORTS,
                      FIELD_FP, FIELD_FQ, CURVE_VESTA, CURVE_PALLAS)
from . import arithmetic

__all__ = ["Context", "BzError", "lib_path", "load_library", "build_library", "EXPORTS", "arithmetic",
           "FIELD_FP", "FIELD_FQ", "CURVE_VESTA", "CURVE_PALLAS"]
