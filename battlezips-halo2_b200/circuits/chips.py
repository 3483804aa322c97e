"""Mirrors of the reference's own chips, gate for gate and region for region (constraint and region names included, because
the reference's MockProver tests quote them):
  bitify     R:src/chips/bitify.rs:44-150 (Num2BitsChip), :153-247 (Bits2NumChip)
  placement  R:src/chips/placement.rs:102-257 (configure), :259-361 (synthesize), :380-419 (trace), :446-568 (state)
  transpose  R:src/chips/transpose.rs:42-137"""
from ..plonk.circuit import Constant
from .fixed_bases import lagrange_interpolate

P = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001
BOARD_SIZE = 100            # R:src/utils/board.rs:12


def bits_of(value, count=BOARD_SIZE):
    """BinaryValue::bitfield (R:src/utils/binary.rs:81-88): little-endian bits."""
    return [(value >> i) & 1 for i in range(count)]


# ---- bitify -----------------------------------------------------------------------------------------------------
def bitify_configure(cs, name, bits, lc1, e2):
    selector = cs.selector()
    one = Constant(1)
    bit = cs.query_advice(bits, 0)
    e2_exp, e2_next = cs.query_advice(e2, 0), cs.query_advice(e2, 1)
    lc1_exp, lc1_next = cs.query_advice(lc1, 0), cs.query_advice(lc1, 1)
    s = cs.query_fixed(selector)
    cs.create_gate(name, [("Constrain bit is boolean", s * (bit * (one - bit))),
                          ("Start from 1, doubling", s * (e2_exp + e2_exp - e2_next)),
                          ("If bit is 1, e2 added to sum", s * (bit * e2_exp + lc1_exp - lc1_next))])
    return {"bits": bits, "lc1": lc1, "e2": e2, "selector": selector}


def num2bits_synthesize(lay, cfg, value_cell, bits):
    """Num2BitsChip::synthesize: region "num2bits"; returns the bit cells; lc1 after the last bit equals the value."""
    def body(region):
        lc1 = region.assign_advice_from_constant(cfg["lc1"], 0, 0)
        e2 = region.assign_advice_from_constant(cfg["e2"], 0, 1)
        cells = []
        for i, b in enumerate(bits):
            region.enable_selector(cfg["selector"], i)
            bit = region.assign_advice(cfg["bits"], i, b)
            cells.append(bit)
            lc1 = region.assign_advice(cfg["lc1"], i + 1, lc1.value + bit.value * e2.value)
            e2 = region.assign_advice(cfg["e2"], i + 1, e2.value + e2.value)
        region.constrain_equal(value_cell, lc1)
        return cells
    return lay.assign_region("num2bits", body)


def bits2num_synthesize(lay, cfg, bit_cells):
    """Bits2NumChip::synthesize: region "bits2num"; copies the bits in, returns the composed value cell."""
    def body(region):
        lc1 = region.assign_advice_from_constant(cfg["lc1"], 0, 0)
        e2 = region.assign_advice_from_constant(cfg["e2"], 0, 1)
        for i, src in enumerate(bit_cells):
            region.enable_selector(cfg["selector"], i)
            bit = region.copy_advice(src, cfg["bits"], i)
            lc1 = region.assign_advice(cfg["lc1"], i + 1, lc1.value + bit.value * e2.value)
            e2 = region.assign_advice(cfg["e2"], i + 1, e2.value + e2.value)
        return lc1
    return lay.assign_region("bits2num", body)


# ---- placement --------------------------------------------------------------------------------------------------
def placement_configure(cs, S, bits, bit_sum, full_window_sum):
    s_input, s_sum_bits, s_adjacency, s_permute, s_constrain = [cs.selector() for _ in range(5)]
    A = cs.query_advice
    cs.create_gate("sum inputted H, V bits", [("h + v = sum", cs.query_fixed(s_input) * (A(bits, 0) - (A(bit_sum, 0) + A(full_window_sum, 0))))])
    cs.create_gate("placement bit count", [("Running Sum: Bits", cs.query_fixed(s_sum_bits) * (A(bits, 0) + A(bit_sum, -1) - A(bit_sum, 0)))])
    bit_count = A(bits, 0)
    for i in range(1, S):
        bit_count = bit_count + A(bits, i)
    prev_full, full = A(full_window_sum, -1), A(full_window_sum, 0)
    coeffs = lagrange_interpolate(list(range(S + 1)), [1 if i == S else 0 for i in range(S + 1)])

    def exp_pow(base, power):
        if power == 0:
            return Constant(1)
        e = base
        for _ in range(2, power + 1):
            e = e * base
        return e
    incr = Constant(0)
    for i, c in enumerate(coeffs):
        incr = incr + Constant(c) * exp_pow(bit_count, i)
    cs.create_gate("adjacency bit count", [("Full Window Running Sum", cs.query_fixed(s_adjacency) * (full - prev_full - incr))])
    cs.create_gate("permute adjaceny bit count", [("Premute Full Window Running Sum", cs.query_fixed(s_permute) * (A(full_window_sum, -1) - A(full_window_sum, 0)))])
    q = cs.query_fixed(s_constrain)
    cs.create_gate("running sum constraints", [("Placed ship of correct length", q * (A(bit_sum, 0) - Constant(S))),
                                               ("One full bit window", q * (A(full_window_sum, 0) - Constant(1)))])
    return {"S": S, "bits": bits, "bit_sum": bit_sum, "full_window_sum": full_window_sum, "s_input": s_input, "s_sum_bits": s_sum_bits,
            "s_adjacency": s_adjacency, "s_permute": s_permute, "s_constrain": s_constrain}


def placement_trace(bits, S):
    """compute_placement_trace (R:src/chips/placement.rs:380-419)."""
    bit_sum, acc = [], 0
    for b in bits:
        acc += b
        bit_sum.append(acc)
    inc = lambda off: 1 if sum(bits[off:off + S]) == S else 0
    full = [inc(0)]
    for i in range(1, len(bits)):
        full.append(full[-1] if i % 10 + S > 10 else full[-1] + inc(i))
    return bit_sum, full


def placement_synthesize(lay, cfg, ship_value, horizontal, vertical):
    """PlacementChip::synthesize: three regions."""
    S = cfg["S"]
    bits = bits_of(ship_value)
    trace = placement_trace(bits, S)

    def load_bits(region):
        out = []
        for i in range(BOARD_SIZE):
            region.enable_selector(cfg["s_input"], i)
            region.copy_advice(horizontal[i], cfg["bit_sum"], i)
            region.copy_advice(vertical[i], cfg["full_window_sum"], i)
            out.append(region.assign_advice(cfg["bits"], i, bits[i]))
        return out
    assigned = lay.assign_region("permute and collapse bit decompositions", load_bits)

    def sums(region):
        region.assign_advice_from_constant(cfg["bit_sum"], 0, 0)
        region.assign_advice_from_constant(cfg["full_window_sum"], 0, 0)
        for i, bit in enumerate(assigned):
            region.copy_advice(bit, cfg["bits"], i + 1)
        bs = region.assign_advice(cfg["bit_sum"], 1, trace[0][0])
        fw = region.assign_advice(cfg["full_window_sum"], 1, trace[1][0])
        region.enable_selector(cfg["s_sum_bits"], 1)
        region.enable_selector(cfg["s_adjacency"], 1)
        for offset in range(2, BOARD_SIZE + 1):
            adj = offset - 1
            bs = region.assign_advice(cfg["bit_sum"], offset, trace[0][adj])
            fw = region.assign_advice(cfg["full_window_sum"], offset, trace[1][adj])
            region.enable_selector(cfg["s_sum_bits"], offset)
            region.enable_selector(cfg["s_permute"] if adj % 10 + S > 10 else cfg["s_adjacency"], offset)
        return bs, fw
    bs, fw = lay.assign_region("placement running sum trace", sums)

    def constrain(region):
        region.copy_advice(bs, cfg["bit_sum"], 0)
        region.copy_advice(fw, cfg["full_window_sum"], 0)
        region.enable_selector(cfg["s_constrain"], 0)
    lay.assign_region("constrain running sum output", constrain)


# ---- transpose --------------------------------------------------------------------------------------------------
def transpose_configure(cs, permuted_bits, transposed_bits):
    selector = cs.selector()
    one = Constant(1)
    t = Constant(0)
    for c in permuted_bits:
        t = t + cs.query_advice(c, 0)
    trace = cs.query_advice(transposed_bits, 0)
    s = cs.query_fixed(selector)
    cs.create_gate("transpose row constraint", [("Constrain trace value integrity", s * (trace - t)),
                                                ("Constrain transposition of bit", s * ((one - t) * t))])
    return {"permuted_bits": list(permuted_bits), "transposed_bits": transposed_bits, "selector": selector}


def transpose_synthesize(lay, cfg, board_bits, placements):
    def body(region):
        for col in range(10):
            for row in range(BOARD_SIZE):
                t = row % 10 * 10 + row // 10 if col % 2 == 1 else row
                region.copy_advice(placements[col][t], cfg["permuted_bits"][col], row)
        out = []
        for row in range(BOARD_SIZE):
            out.append(region.assign_advice(cfg["transposed_bits"], row, board_bits[row]))
            region.enable_selector(cfg["selector"], row)
        return out
    return lay.assign_region("Transpose ship commitments", body)
