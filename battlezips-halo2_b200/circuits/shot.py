"""Mirror of the reference ShotCircuit (R:src/circuits/shot.rs:22-53, chip R:src/chips/shot.rs:179-354), k = 11.

11 advice columns (10 + `input`), 8 user fixed columns (fixed[0] = constants), 1 table column, 1 instance column
(4 public values), selectors as fixed columns; gate order as in the reference: 2 x num2bits, 19 ECC/range-check
gates (stand-ins, see circuits/common.py), then "boolean hit assertion", "shot running sum row", "constrain shot
running sum output" -- 24 gates, 1 lookup, 13 permutation columns, degree 9."""
import random
from ..plonk.circuit import ConstraintSystem, Constant
from .common import (P, BOARD_SIZE, Layout, num2bits_configure, num2bits_synthesize, ecc_shape_configure,
                     ecc_shape_load_table, ecc_shape_synthesize)

K = 11                     # R:benches/shot.rs:22


def configure():
    cs = ConstraintSystem(P)
    advice = [cs.advice_column() for _ in range(10)]
    for c in advice:
        cs.enable_equality("advice", c)
    inp = cs.advice_column()
    cs.enable_equality("advice", inp)
    fixed = [cs.fixed_column() for _ in range(8)]
    cs.enable_equality("fixed", fixed[0])          # enable_constant
    table = cs.fixed_column()
    instance = cs.instance_column()
    cs.enable_equality("instance", instance)
    selectors = [cs.fixed_column() for _ in range(3)]
    num2bits = [num2bits_configure(cs, advice[5], advice[6], advice[7]) for _ in range(2)]
    ecc = ecc_shape_configure(cs, advice, fixed, table)
    A = lambda i, r=0: cs.query_advice(advice[i], r)
    one = Constant(1)
    cs.create_gate("boolean hit assertion", [cs.query_fixed(selectors[0]) * ((one - A(4)) * A(4))])
    s1 = cs.query_fixed(selectors[1])
    cs.create_gate("shot running sum row", [s1 * (A(6) + A(7, -1) - A(7)),
                                            s1 * (A(5) * A(6) + A(8, -1) - A(8))])
    s2 = cs.query_fixed(selectors[2])
    cs.create_gate("constrain shot running sum output", [s2 * (one - A(6)), s2 * (A(5) - A(7))])
    return cs, {"advice": advice, "input": inp, "fixed": fixed, "table": table, "instance": instance,
                "selectors": selectors, "num2bits": num2bits, "ecc": ecc}


def synthesize(cs, cfg, board_bits, shot_bits, hit, trapdoor, seed=0):
    """R:src/chips/shot.rs:308-354.  board_bits / shot_bits: 100 booleans; hit: 0/1; trapdoor: Fq scalar (int)."""
    rng = random.Random(seed)
    lay = Layout(cs, K)
    a, adv, fx, sel = lay.asg, cfg["advice"], cfg["fixed"], cfg["selectors"]
    ecc_shape_load_table(lay, cfg["ecc"])
    board_state = sum(b << i for i, b in enumerate(board_bits))
    shot_commitment = sum(b << i for i, b in enumerate(shot_bits))
    # commitment stand-in region first so its coordinates are known for load_advice
    # load_advice: advice[4] rows 0..4
    r0 = lay.region(5)
    cells = [("advice", adv[4], r0 + i) for i in range(5)]
    # decompose
    bits_cells = [num2bits_synthesize(lay, cfg["num2bits"][0], fx[0], cells[0], board_bits),
                  num2bits_synthesize(lay, cfg["num2bits"][1], fx[0], cells[3], shot_bits)]
    # running sums (R:src/chips/shot.rs:438-493)
    rr = lay.region(BOARD_SIZE + 1)
    a.assign_advice(adv[7], rr, 0)
    a.assign_advice(adv[8], rr, 0)
    a.copy(("advice", adv[7], rr), lay.constant(fx[0], 0))
    a.copy(("advice", adv[8], rr), lay.constant(fx[0], 0))
    shot_sum = hit_sum = 0
    for i in range(BOARD_SIZE):
        row = rr + i + 1
        a.assign_advice(adv[5], row, board_bits[i]); a.copy(bits_cells[0][i], ("advice", adv[5], row))
        a.assign_advice(adv[6], row, shot_bits[i]); a.copy(bits_cells[1][i], ("advice", adv[6], row))
        shot_sum += shot_bits[i]
        hit_sum += board_bits[i] & shot_bits[i]
        a.assign_advice(adv[7], row, shot_sum)
        a.assign_advice(adv[8], row, hit_sum)
        a.assign_fixed(sel[1], row, 1)
    # running sum output (R:src/chips/shot.rs:495-525)
    ro = lay.region(1)
    a.assign_advice(adv[5], ro, hit); a.copy(cells[4], ("advice", adv[5], ro))
    a.assign_advice(adv[6], ro, shot_sum); a.copy(("advice", adv[7], rr + BOARD_SIZE), ("advice", adv[6], ro))
    a.assign_advice(adv[7], ro, hit_sum); a.copy(("advice", adv[8], rr + BOARD_SIZE), ("advice", adv[7], ro))
    a.assign_fixed(sel[2], ro, 1)
    # board commitment [v]V + [r]R (stand-in arithmetic, same shape)
    cx, cy, commit = ecc_shape_synthesize(lay, cfg["ecc"], board_state, trapdoor, rng)
    for i, v in enumerate((board_state, commit[0], commit[1], shot_commitment, hit)):
        a.assign_advice(adv[4], r0 + i, v)
    a.assign_fixed(sel[0], r0 + 4, 1)
    a.copy(cx, cells[1])
    a.copy(cy, cells[2])
    # public inputs (R:src/chips/shot.rs:349-352)
    a.set_instance(cfg["instance"], [commit[0], commit[1], shot_commitment, hit])
    a.copy(cx, ("instance", cfg["instance"], 0))
    a.copy(cy, ("instance", cfg["instance"], 1))
    a.copy(cells[3], ("instance", cfg["instance"], 2))
    a.copy(cells[4], ("instance", cfg["instance"], 3))
    return a


# board "pattern 1" of the reference tests: R:src/circuits/shot.rs:102-108 / R:benches/board.rs:26-32
PATTERN_1 = [(3, 3, 1), (5, 4, 0), (0, 1, 0), (0, 5, 1), (6, 1, 0)]      # (x, y, vertical)
PATTERN_2 = [(3, 4, 1), (9, 6, 1), (0, 0, 0), (0, 6, 0), (6, 1, 1)]
SHIP_LENGTHS = [5, 4, 3, 3, 2]


def board_bits_from_pattern(pattern):
    bits = [0] * BOARD_SIZE
    for (x, y, vertical), length in zip(pattern, SHIP_LENGTHS):
        for i in range(length):
            cx, cy = (x, y + i) if vertical else (x + i, y)
            assert 0 <= cx < 10 and 0 <= cy < 10
            bits[cy * 10 + cx] = 1
    return bits


def shot_circuit(index=0, seed=0):
    """The synthetic Shot job #index: board pattern index mod 2, shot cell
    (index mod 10, (index/10) mod 10), correct hit bit.  Returns (cs, cfg, assignment)."""
    cs, cfg = configure()
    pattern = PATTERN_1 if index % 2 == 0 else PATTERN_2
    board_bits = board_bits_from_pattern(pattern)
    sx, sy = index % 10, (index // 10) % 10
    shot_bits = [0] * BOARD_SIZE
    shot_bits[sy * 10 + sx] = 1
    hit = board_bits[sy * 10 + sx]
    trapdoor = random.Random(1000 + seed + index).randrange(1 << 254)
    asg = synthesize(cs, cfg, board_bits, shot_bits, hit, trapdoor, seed=seed + index)
    return cs, cfg, asg
