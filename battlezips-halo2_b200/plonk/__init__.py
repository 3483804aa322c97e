"""Host-side mirror of the `halo2_proofs::plonk` interface used by the reference
(/root/reference/benches/shot.rs:11-17,58-71): ConstraintSystem / Expression, keygen, create_proof."""
from .circuit import (ConstraintSystem, Expression, Constant, Advice, Fixed, Instance, Rotation, Assignment)
