"""Host-side mirrors of the reference circuits (R:src/circuits/shot.rs, R:src/circuits/board.rs) on the Python
ConstraintSystem mirror.  The repo's own chips (bitify, running sums, orientation, boolean checks) are restated
gate-for-gate; the 19 halo2_gadgets ECC / range-check gates that `PedersenCommitmentChip::configure` pulls in
(R:src/chips/pedersen.rs:49-62) are SHAPE-EQUIVALENT stand-ins (same count, degrees up to 9, rotations, one
degree-3 lookup against the 1024-row table) -- see SURVEY App. C and DESIGN.md "Circuits"."""
from .shot import shot_circuit
from .board import board_circuit, board_circuit_scaled
