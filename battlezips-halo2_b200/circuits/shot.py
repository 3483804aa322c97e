"""Mirror of the reference ShotCircuit (R:src/circuits/shot.rs:22-53; chip R:src/chips/shot.rs:179-297 configure, :308-536
synthesize), k = 11: 10 + 1 advice columns, 8 fixed (fixed[0] = constants, all 8 = the Lagrange-coefficient columns of the
ECC chip), 1 table column, 1 instance column (commitment x, y, shot, hit), 3 own selectors; gates in the reference's order:
2 x num2bits, 19 of halo2_gadgets (circuits/gadgets.py), then #21 "boolean hit assertion", #22 "shot running sum row",
#23 "constrain shot running sum output"; regions 0 load, 1-2 num2bits, 3 running sum, 4 output, 5 table, 6-11 ECC, 12 final
addition -- the numbering the reference's MockProver tests quote (R:src/circuits/shot.rs:297-332, 681-690)."""
import random
from ..plonk.circuit import ConstraintSystem, Constant, Layouter, compress_selectors
from .chips import P, BOARD_SIZE, bits_of, bitify_configure, num2bits_synthesize
from .gadgets import PedersenCommitmentChip
from .fixed_bases import pedersen_commit, Q

K = 11                     # R:benches/shot.rs:22


def configure():
    cs = ConstraintSystem(P)
    advice = [cs.advice_column() for _ in range(10)]
    for c in advice:
        cs.enable_equality("advice", c)
    inp = cs.advice_column()
    cs.enable_equality("advice", inp)
    fixed = [cs.fixed_column() for _ in range(8)]
    cs.enable_constant(fixed[0])
    table_idx = cs.lookup_table_column()
    instance = cs.instance_column()
    cs.enable_equality("instance", instance)
    selectors = [cs.selector() for _ in range(3)]
    num2bits = [bitify_configure(cs, "num2bits", advice[5], advice[6], advice[7]) for _ in range(2)]
    pedersen = PedersenCommitmentChip(cs, advice, fixed, table_idx)
    A = lambda i, r=0: cs.query_advice(advice[i], r)
    one = Constant(1)
    cs.create_gate("boolean hit assertion", [("asserted hit value is boolean", cs.query_fixed(selectors[0]) * ((one - A(4)) * A(4)))])
    hit_bit, shot_bit, shot_sum, hit_sum = A(5), A(6), A(7), A(8)
    s1 = cs.query_fixed(selectors[1])
    cs.create_gate("shot running sum row", [("running sum of flipped bits in shot", s1 * (shot_bit + A(7, -1) - shot_sum)),
                                            ("running sum of hits against board", s1 * (hit_bit * shot_bit + A(8, -1) - hit_sum))])
    s2 = cs.query_fixed(selectors[2])
    cs.create_gate("constrain shot running sum output", [("Shot only fires at one board cell", s2 * (one - A(6))),
                                                         ("Public hit assertion matches private witness", s2 * (A(5) - A(7)))])
    return cs, {"advice": advice, "input": inp, "fixed": fixed, "table": table_idx, "instance": instance,
                "selectors": selectors, "num2bits": num2bits, "pedersen": pedersen}


def synthesize(cs, cfg, board_state, trapdoor, shot, hit, public=None, k=K):
    """ShotChip::synthesize (R:src/chips/shot.rs:308-354).  board_state / shot: 100-bit integers, hit: the asserted value
    (any integer: the negative tests assert 2), trapdoor: a pallas::Scalar.  public: the instance column (default: the honest
    [commitment x, y, shot, hit])."""
    lay = Layouter(cs, k)
    adv, sel = cfg["advice"], cfg["selectors"]
    commitment = pedersen_commit(board_state, trapdoor)
    board_bits, shot_bits = bits_of(board_state), bits_of(shot)
    shot_trace, hit_trace, s_acc, h_acc = [], [], 0, 0            # compute_shot_trace (R:src/chips/shot.rs:28-51)
    for i in range(BOARD_SIZE):
        s_acc += shot_bits[i]
        h_acc += board_bits[i] & shot_bits[i]
        shot_trace.append(s_acc); hit_trace.append(h_acc)

    def load(region):
        cells = [region.assign_advice(adv[4], i, v) for i, v in enumerate((board_state, commitment[0], commitment[1], shot, hit))]
        region.enable_selector(sel[0], 4)
        return cells
    inputs = lay.assign_region("load private ShotChip advice values", load)
    bits = [num2bits_synthesize(lay, cfg["num2bits"][0], inputs[0], board_bits),
            num2bits_synthesize(lay, cfg["num2bits"][1], inputs[3], shot_bits)]

    def running(region):
        ss = region.assign_advice_from_constant(adv[7], 0, 0)
        hs = region.assign_advice_from_constant(adv[8], 0, 0)
        for i in range(BOARD_SIZE):
            region.copy_advice(bits[0][i], adv[5], i + 1)
            region.copy_advice(bits[1][i], adv[6], i + 1)
            ss = region.assign_advice(adv[7], i + 1, shot_trace[i])
            hs = region.assign_advice(adv[8], i + 1, hit_trace[i])
            region.enable_selector(sel[1], i + 1)
        return ss, hs
    ss, hs = lay.assign_region("shot running sum", running)

    def output(region):
        region.copy_advice(inputs[4], adv[5], 0)
        region.copy_advice(ss, adv[6], 0)
        region.copy_advice(hs, adv[7], 0)
        region.enable_selector(sel[2], 0)
    lay.assign_region("shot running sum output checks", output)
    point = cfg["pedersen"].synthesize(lay, inputs[0], trapdoor)
    assert point[2] == commitment
    inst = cfg["instance"]
    lay.constrain_instance(point[0], inst, 0)
    lay.constrain_instance(point[1], inst, 1)
    lay.constrain_instance(inputs[3], inst, 2)
    lay.constrain_instance(inputs[4], inst, 3)
    lay.asg.set_instance(inst, list(public) if public is not None else [commitment[0], commitment[1], shot, hit])
    return lay.asg


# board "pattern 1" / "pattern 2" of the reference tests: R:src/circuits/shot.rs:102-108, 143-149 (x, y, z = vertical)
PATTERN_1 = [(3, 3, 1), (5, 4, 0), (0, 1, 0), (0, 5, 1), (6, 1, 0)]
PATTERN_2 = [(3, 4, 0), (9, 6, 1), (0, 0, 0), (0, 6, 0), (6, 1, 1)]
SHIP_LENGTHS = [5, 4, 3, 3, 2]


def ship_coordinates(x, y, z, length, transpose):
    """Ship::coordinates (R:src/utils/ship.rs:139-154)."""
    out = []
    for i in range(length):
        x_i, y_i = (x, y + i) if z else (x + i, y)
        out.append((x_i * 10 + y_i) if (transpose and z) else (x_i + y_i * 10))
    return out


def board_state_from_pattern(pattern):
    """Board::state with default witness options (R:src/utils/board.rs:75-96): the 100-bit board integer."""
    state = 0
    for (x, y, z), length in zip(pattern, SHIP_LENGTHS):
        for c in ship_coordinates(x, y, z, length, False):
            state |= 1 << c
    return state


def serialize_shot(x, y):
    """R:src/utils/shot.rs:12-19"""
    return 1 << (y * 10 + x)


def shot_circuit(index=0, seed=0, compress=True):
    """The synthetic Shot job #index: board pattern index mod 2, shot cell (index mod 10, (index / 10) mod 10),
    correct hit bit, trapdoor from a seeded generator.  Returns (cs, cfg, assignment) -- after selector compression, i.e. what
    keygen sees, unless compress=False (the form MockProver reports regions and selectors on)."""
    cs, cfg = configure()
    board = board_state_from_pattern(PATTERN_1 if index % 2 == 0 else PATTERN_2)
    shot = serialize_shot(index % 10, (index // 10) % 10)
    hit = 1 if board & shot else 0
    trapdoor = random.Random(1000 + seed + index).randrange(Q)
    asg = synthesize(cs, cfg, board, trapdoor, shot, hit)
    if compress:
        cs, asg = compress_selectors(cs, asg)
    return cs, cfg, asg
