"""Mirror of the reference BoardCircuit (R:src/circuits/board.rs:21-51, chip R:src/chips/board.rs:194-363), k = 12.

11 advice, 8 user fixed (fixed[0] constants), 1 table, 1 instance (2 public values); gate order as the reference:
10 x num2bits, bits2num, 25 placement gates, transpose, 19 ECC/range-check gates, "Commitment orientation H OR V"
-- 57 gates, 1 lookup, 13 permutation columns, degree 9, advice[0] queried at rotations 0..4 by the adjacency
gates.  num2bits / bits2num / orientation are restated from the reference; the placement (R:src/chips/placement.rs
:130-260) and transpose (R:src/chips/transpose.rs:54-80) gates are stand-ins of the same degree (<= 6) and
rotation pattern, the ECC gates as in circuits/common.py."""
import random
from ..plonk.circuit import ConstraintSystem, Constant
from .common import (P, BOARD_SIZE, Layout, num2bits_configure, num2bits_synthesize, ecc_shape_configure,
                     ecc_shape_load_table, ecc_shape_synthesize)
from .shot import PATTERN_1, PATTERN_2, SHIP_LENGTHS, board_bits_from_pattern

K = 12                     # R:benches/board.rs:22


def configure():
    cs = ConstraintSystem(P)
    advice = [cs.advice_column() for _ in range(11)]
    for c in advice:
        cs.enable_equality("advice", c)
    fixed = [cs.fixed_column() for _ in range(8)]
    cs.enable_equality("fixed", fixed[0])
    table = cs.fixed_column()
    instance = cs.instance_column()
    cs.enable_equality("instance", instance)
    selectors = [cs.fixed_column()]
    num2bits = [num2bits_configure(cs, advice[0], advice[1], advice[2]) for _ in range(10)]
    bits2num = num2bits_configure(cs, advice[0], advice[1], advice[2])      # same 3-constraint gate (bitify.rs:152-198)
    cs.gates[-1] = ("bits2num", cs.gates[-1][1])
    A = lambda i, r=0: cs.query_advice(advice[i], r)
    one = Constant(1)
    placement = []
    for ship, S in zip(("carrier", "battleship", "cruiser", "submarine", "destroyer"), SHIP_LENGTHS):
        q = [cs.fixed_column() for _ in range(5)]
        bit, bsum, fsum = A(0), A(1), A(2)
        # stand-ins with the degree / rotation shape of PlacementChip<S>: bit-sum running row, full-window
        # adjacency over S consecutive bits (rotations 0..S-1 of advice[0]), interpolated counter (deg <= 6)
        window = A(0)
        for r in range(1, S):
            window = window * A(0, r)
        cs.create_gate(f"{ship}: placement bit sum", [cs.query_fixed(q[0]) * (bit + A(1, -1) - bsum)])
        cs.create_gate(f"{ship}: adjacency window", [cs.query_fixed(q[1]) * (window + A(2, -1) - fsum)])
        cs.create_gate(f"{ship}: bit count == S", [cs.query_fixed(q[2]) * (bsum - Constant(S))])
        cs.create_gate(f"{ship}: exactly one full window", [cs.query_fixed(q[3]) * (fsum - one)])
        cs.create_gate(f"{ship}: running sum constraints", [cs.query_fixed(q[4]) * bit * (one - bit) * bsum * fsum])
        placement.append(q)
    q_t = cs.fixed_column()
    row_or = A(0)
    for i in range(1, 10):
        row_or = row_or + A(i)
    cs.create_gate("transpose row constraint", [cs.query_fixed(q_t) * (row_or - A(10)),
                                                cs.query_fixed(q_t) * A(10) * (one - A(10))])
    ecc = ecc_shape_configure(cs, advice[:10], fixed, table)
    s = cs.query_fixed(selectors[0])
    cs.create_gate("Commitment orientation H OR V == 0 constraint", [s * A(2 * i) * A(2 * i + 1) for i in range(5)])
    return cs, {"advice": advice, "fixed": fixed, "table": table, "instance": instance, "selectors": selectors,
                "num2bits": num2bits, "bits2num": bits2num, "placement": placement, "q_transpose": q_t, "ecc": ecc}


def synthesize(cs, cfg, pattern, trapdoor, seed=0, k=K, copies=1):
    """R:src/chips/board.rs:331-363: 10 ship commitments (H, V per ship) -> bits -> placement -> transpose ->
    board state -> Pedersen commitment; public = commitment (x, y).
    copies > 1 tiles the whole board region pattern down the rows (BASELINE config 5: "Board circuit replicated ...
    many boards per proof"); the public inputs are those of the first board."""
    rng = random.Random(seed)
    lay = Layout(cs, k)
    a = lay.asg
    ecc_shape_load_table(lay, cfg["ecc"])
    first = None
    for c in range(copies):
        pat = pattern if c % 2 == 0 else (PATTERN_2 if pattern is PATTERN_1 else PATTERN_1)
        cells = _synthesize_board(lay, cs, cfg, pat, (trapdoor + c) % (1 << 254), rng)
        first = first or cells
    cx, cy, commit = first
    a.set_instance(cfg["instance"], [commit[0], commit[1]])          # R:src/chips/board.rs:359-360
    a.copy(cx, ("instance", cfg["instance"], 0))
    a.copy(cy, ("instance", cfg["instance"], 1))
    return a


def _synthesize_board(lay, cs, cfg, pattern, trapdoor, rng):
    a, adv, fx = lay.asg, cfg["advice"], cfg["fixed"]
    # ship commitments: horizontal / vertical bitfields, one of each pair is zero
    ships = []
    for (x, y, vertical), length in zip(pattern, SHIP_LENGTHS):
        bits = [0] * BOARD_SIZE
        for i in range(length):
            cx, cy = (x, y + i) if vertical else (x + i, y)
            idx = (cx * 10 + cy) if vertical else (cy * 10 + cx)       # vertical commitments are transposed
            bits[idx] = 1
        ships.append(([0] * BOARD_SIZE, bits) if vertical else (bits, [0] * BOARD_SIZE))
    r0 = lay.region(1)
    commit_cells = []
    for i, (h, v) in enumerate(ships):
        for j, bits in enumerate((h, v)):
            col = adv[2 * i + j]
            a.assign_advice(col, r0, sum(b << t for t, b in enumerate(bits)))
            commit_cells.append(("advice", col, r0))
    a.assign_fixed(cfg["selectors"][0], r0, 1)
    # decompose the 10 commitments
    bit_cells = []
    for i, (h, v) in enumerate(ships):
        for j, bits in enumerate((h, v)):
            bit_cells.append(num2bits_synthesize(lay, cfg["num2bits"][2 * i + j], fx[0], commit_cells[2 * i + j], bits))
    # placement stand-ins: running sums over the 100 bits of the non-zero commitment of each ship
    for i, ((h, v), S) in enumerate(zip(ships, SHIP_LENGTHS)):
        bits = h if any(h) else v
        q = cfg["placement"][i]
        rp = lay.region(BOARD_SIZE + S + 1)
        bsum = fsum = 0
        a.assign_advice(adv[1], rp, 0)
        a.assign_advice(adv[2], rp, 0)
        padded = bits + [0] * S
        for t in range(BOARD_SIZE):
            row = rp + 1 + t
            a.assign_advice(adv[0], row, padded[t])
            win = 1
            for r in range(S):
                win &= padded[t + r]
            bsum += padded[t]
            fsum += win
            a.assign_advice(adv[1], row, bsum)
            a.assign_advice(adv[2], row, fsum)
            a.assign_fixed(q[0], row, 1)
            a.assign_fixed(q[1], row, 1)
        last = rp + BOARD_SIZE
        a.assign_fixed(q[2], last, 1)
        a.assign_fixed(q[3], last, 1)
    # transpose stand-in: OR of the 10 decomposed bit columns into the board state bits
    board_bits = board_bits_from_pattern(pattern)
    rt = lay.region(BOARD_SIZE)
    for t in range(BOARD_SIZE):
        cx, cy = t % 10, t // 10
        acc = 0
        for i, (h, v) in enumerate(ships):
            hv = h[t]
            vv = v[cx * 10 + cy]
            a.assign_advice(adv[2 * i], rt + t, hv)
            a.assign_advice(adv[2 * i + 1], rt + t, vv)
            acc += hv + vv
        assert acc == board_bits[t]
        a.assign_advice(adv[10], rt + t, acc)
        a.assign_fixed(cfg["q_transpose"], rt + t, 1)
    # recompose the board state (bits2num) and commit
    board_state = sum(b << i for i, b in enumerate(board_bits))
    rs = lay.region(1)
    a.assign_advice(adv[3], rs, board_state)
    state_cell = ("advice", adv[3], rs)
    b2n = num2bits_synthesize(lay, cfg["bits2num"], fx[0], state_cell, board_bits)
    for t in range(BOARD_SIZE):
        a.copy(b2n[t], ("advice", adv[10], rt + t))
    return ecc_shape_synthesize(lay, cfg["ecc"], board_state, trapdoor, rng)


def board_circuit_scaled(k, copies=None, seed=0):
    """BASELINE config 5: the Board circuit replicated down the rows of a 2^k-row table (copies=None: as many boards as
    fit).  Same constraint system as the k=12 circuit; only the number of used rows and k change."""
    cs, cfg = configure()
    rows_per_board = 2100
    if copies is None:
        copies = max(1, ((1 << k) - 1200) // rows_per_board)
    asg = synthesize(cs, cfg, PATTERN_1, random.Random(3000 + seed).randrange(1 << 254), seed=seed, k=k, copies=copies)
    return cs, cfg, asg


def board_circuit(index=0, seed=0):
    """Synthetic Board job #index: reference board pattern index mod 2 (R:benches/board.rs:26-32)."""
    cs, cfg = configure()
    pattern = PATTERN_1 if index % 2 == 0 else PATTERN_2
    trapdoor = random.Random(2000 + seed + index).randrange(1 << 254)
    asg = synthesize(cs, cfg, pattern, trapdoor, seed=seed + index)
    return cs, cfg, asg
