"""Host-side mirrors of the reference circuits (R:src/circuits/shot.rs, R:src/circuits/board.rs) on the Python
ConstraintSystem / Layouter mirror: the reference's own chips (chips.py) and the halo2_gadgets 0.2.0 ECC and range-check
chips its Pedersen commitment instantiates (gadgets.py, fixed_bases.py) are restated gate for gate and region for region;
tests/test_circuit_mirrors.py replays the reference's MockProver tests (positive and negative) on them."""
from .shot import shot_circuit
from .board import board_circuit, board_circuit_scaled
