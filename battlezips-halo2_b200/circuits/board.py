"""Mirror of the reference BoardCircuit (R:src/circuits/board.rs:21-51; chip R:src/chips/board.rs:194-321 configure, :331-499
synthesize), k = 12: 11 advice, 8 fixed (fixed[0] constants), 1 table, 1 instance (commitment x, y), 1 own selector; gates in
the reference's order: #0-9 num2bits, #10 bits2num, #11-35 placement (5 ships x 5 gates), #36 transpose, #37-55 halo2_gadgets,
#56 "Commitment orientation H OR V == 0 constraint"; regions 0 load, 1-10 num2bits, 11-25 placement (3 per ship), 26
transpose, 27 bits2num, 28 table, 29-34 ECC, 35 final addition (R:src/circuits/board.rs:198-228, 867)."""
import random
from ..plonk.circuit import ConstraintSystem, Layouter, compress_selectors
from .chips import (P, BOARD_SIZE, bits_of, bitify_configure, num2bits_synthesize, bits2num_synthesize, placement_configure,
                    placement_synthesize, transpose_configure, transpose_synthesize)
from .gadgets import PedersenCommitmentChip
from .fixed_bases import pedersen_commit, Q
from .shot import PATTERN_1, PATTERN_2, SHIP_LENGTHS, ship_coordinates, board_state_from_pattern

K = 12                     # R:benches/board.rs:22
SHIP_NAMES = ["Aircraft Carrier", "Battleship", "Cruiser", "Submarine", "Destroyer"]


def configure():
    cs = ConstraintSystem(P)
    advice = [cs.advice_column() for _ in range(11)]
    for c in advice:
        cs.enable_equality("advice", c)
    fixed = [cs.fixed_column() for _ in range(8)]
    cs.enable_constant(fixed[0])
    table_idx = cs.lookup_table_column()
    instance = cs.instance_column()
    cs.enable_equality("instance", instance)
    selectors = [cs.selector()]
    num2bits = [bitify_configure(cs, "num2bits", advice[0], advice[1], advice[2]) for _ in range(10)]
    bits2num = bitify_configure(cs, "bits2num", advice[0], advice[1], advice[2])
    placement = [placement_configure(cs, S, advice[0], advice[1], advice[2]) for S in SHIP_LENGTHS]
    transpose = transpose_configure(cs, advice[:10], advice[10])
    pedersen = PedersenCommitmentChip(cs, advice[:10], fixed, table_idx)
    s = cs.query_fixed(selectors[0])
    com = [cs.query_advice(advice[i], 0) for i in range(10)]
    cs.create_gate("Commitment orientation H OR V == 0 constraint",
                   [(f"{name} H OR V == 0", s * (com[2 * i] * com[2 * i + 1])) for i, name in enumerate(SHIP_NAMES)])
    return cs, {"advice": advice, "fixed": fixed, "table": table_idx, "instance": instance, "selectors": selectors,
                "num2bits": num2bits, "bits2num": bits2num, "placement": placement, "transpose": transpose, "pedersen": pedersen}


# ---- witness options of the reference's negative tests (R:src/utils/ship.rs:189-331) ------------------------------------
def ship_witness(x, y, z, length, option="Default"):
    """Ship::witness -> (horizontal, vertical) 100-bit commitments."""
    coords = ship_coordinates(x, y, z, length, True)
    placement = sum(1 << c for c in coords)
    hv = [0, placement] if z else [placement, 0]
    t = 1 if z else 0
    if option == "DualPlacement":
        hv[1 - t] |= 1 << coords[0]
        hv[t] &= ~(1 << coords[0])
    elif option == "Nonconsecutive":
        hv[t] &= ~(1 << coords[-1])
        hv[t] |= 1 << (coords[-1] + 1)
    elif option == "ExtraBit":
        hv[t] |= 1
    elif option == "Oversized":
        hv[t] |= 1 << (coords[-1] + 1)
    elif option == "Undersized":
        hv[t] &= ~(1 << coords[-1])
    else:
        assert option == "Default", option
    return hv


def board_witness(pattern, options=("Default",) * 5):
    """Board::witness and Board::state under the given witness options (R:src/utils/board.rs:75-117)."""
    commitments, state = [], 0
    for (x, y, z), length, opt in zip(pattern, SHIP_LENGTHS, options):
        h, v = ship_witness(x, y, z, length, opt)
        commitments += [h, v]
        for j in range(BOARD_SIZE):
            if (h >> j) & 1:
                state |= 1 << j
            if (v >> j) & 1:
                state |= 1 << (j % 10 * 10 + j // 10)
    return commitments, state


def _synthesize_board(lay, cfg, ship_commitments, board_state, trapdoor, load_table):
    """BoardChip::synthesize (R:src/chips/board.rs:331-363) up to the commitment point."""
    adv = cfg["advice"]
    ships = []
    for i in range(5):
        h, v = ship_commitments[2 * i], ship_commitments[2 * i + 1]
        assert h & v == 0, f"Cannot zip together bit (ship {i})"            # BinaryValue::zip panics (R:src/utils/binary.rs:97-108)
        ships.append(h | v)

    def load(region):
        cells = [region.assign_advice(adv[i], 0, ship_commitments[i]) for i in range(10)]
        region.enable_selector(cfg["selectors"][0], 0)
        return cells
    assigned = lay.assign_region("load ship placements", load)
    placements = [num2bits_synthesize(lay, cfg["num2bits"][i], assigned[i], bits_of(ship_commitments[i])) for i in range(10)]
    for i in range(5):
        placement_synthesize(lay, cfg["placement"][i], ships[i], placements[2 * i], placements[2 * i + 1])
    transposed_bits = transpose_synthesize(lay, cfg["transpose"], bits_of(board_state), placements)
    transposed = bits2num_synthesize(lay, cfg["bits2num"], transposed_bits)
    return cfg["pedersen"].synthesize(lay, transposed, trapdoor, load_table=load_table)


def synthesize(cs, cfg, ship_commitments, board_state, trapdoor, public=None, k=K):
    lay = Layouter(cs, k)
    point = _synthesize_board(lay, cfg, ship_commitments, board_state, trapdoor, True)
    inst = cfg["instance"]
    lay.constrain_instance(point[0], inst, 0)
    lay.constrain_instance(point[1], inst, 1)
    lay.asg.set_instance(inst, list(public) if public is not None else [point[0].value, point[1].value])
    return lay.asg


def board_circuit(index=0, seed=0, compress=True):
    """Synthetic Board job #index: reference board pattern index mod 2 (R:benches/board.rs:26-32)."""
    cs, cfg = configure()
    commitments, state = board_witness(PATTERN_1 if index % 2 == 0 else PATTERN_2)
    assert state == board_state_from_pattern(PATTERN_1 if index % 2 == 0 else PATTERN_2)
    trapdoor = random.Random(2000 + seed + index).randrange(Q)
    asg = synthesize(cs, cfg, commitments, state, trapdoor)
    if compress:
        cs, asg = compress_selectors(cs, asg)
    return cs, cfg, asg


ROWS_PER_BOARD = 2400       # rows one BoardChip::synthesize adds to the tallest columns (advice[0..2]) under the simple floor planner


def board_circuit_scaled(k, copies=None, seed=0, compress=True):
    """BASELINE config 5: the Board circuit replicated down the rows of a 2^k-row table (copies = None: as many boards as fit).
    Same constraint system as the k = 12 circuit; every copy is a complete BoardChip::synthesize (alternating board patterns,
    its own trapdoor); the lookup table is loaded once; the public inputs are the commitment of the first board."""
    cs, cfg = configure()
    lay = Layouter(cs, k)
    rng = random.Random(3000 + seed)
    if copies is None:
        copies = max(1, lay.asg.usable_rows // ROWS_PER_BOARD)
    first = None
    for c in range(copies):
        commitments, state = board_witness(PATTERN_1 if c % 2 == 0 else PATTERN_2)
        point = _synthesize_board(lay, cfg, commitments, state, rng.randrange(Q), load_table=(c == 0))
        first = first or point
    inst = cfg["instance"]
    lay.constrain_instance(first[0], inst, 0)
    lay.constrain_instance(first[1], inst, 1)
    lay.asg.set_instance(inst, [first[0].value, first[1].value])
    if compress:
        cs2, asg2 = compress_selectors(cs, lay.asg)
        return cs2, cfg, asg2
    return cs, cfg, lay.asg
