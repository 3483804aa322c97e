"""Fixed bases of the board commitment [v]V + [r]R (R:src/utils/pedersen.rs:18-27, R:src/chips/pedersen.rs:104-134) and the
per-window constants halo2_gadgets' fixed-base scalar multiplication loads into fixed columns
(U: halo2_gadgets 0.2.0 src/ecc/chip/constants.rs `compute_window_table`, `compute_lagrange_coeffs`, `find_zs_and_us`).

V, R = hash_to_curve("battlezips:hash2curve")(b"v" / b"r"): the two generator constants of the reference
(R:src/utils/constants/fixed_bases/board_commit_v.rs:5-14, board_commit_r.rs:5-14; reproduced on the device by
bz_hash_to_curve, tests/test_gpu_params.py).  The per-window z are circuit constants
the reference hard-codes as well (board_commit_v.rs:17-26, board_commit_r.rs:17-26; searching the smallest admissible z is a
keygen-time job of minutes in Python): they are carried as data and CHECKED at construction (y + z square, z - y non-square
for all 8 points of the window).  Window tables, Lagrange coefficients and all 85 x 8 u values are computed here;
tests/test_circuit_mirrors.py compares them with the reference's tables (tests/golden/pallas_fixed_base_kats.json)."""
from functools import lru_cache
from . import pallas as E

P, Q = E.P, E.Q
NUM_WINDOWS = 85                     # R:src/utils/constants.rs:4   ceil(255 / 3)
H = 8                                # 2^FIXED_BASE_WINDOW_SIZE
WINDOW_BITS = 3

BOARD_COMMIT_V = (int.from_bytes(bytes.fromhex("a42c69a69962af0ad78513ae5c657dbda3678426f9c33faa5821c416d242251e"), "little"),
                  int.from_bytes(bytes.fromhex("b2140f88aad7a9372f47ba7483a005e718d3ff8cbcf1260af88693034ac9c532"), "little"))
BOARD_COMMIT_R = (int.from_bytes(bytes.fromhex("77520e73bb956edf555cc6e590cb93160cc48b1dc8d2b7cfd7c3a9a16f2c331c"), "little"),
                  int.from_bytes(bytes.fromhex("b8507984802396043f8952a99cffc5cc950d19af64a9d675e5d08a2e7490980f"), "little"))


def lagrange_interpolate(points, evals):
    """halo2_proofs::arithmetic::lagrange_interpolate: coefficients (low degree first) of the polynomial through the points."""
    n = len(points)
    out = [0] * n
    for j in range(n):
        # numerator polynomial prod_{m != j} (X - x_m), denominator prod (x_j - x_m)
        num, den = [1], 1
        for m in range(n):
            if m == j:
                continue
            num = [(a - points[m] * b) % P for a, b in zip([0] + num, num + [0])]
            den = den * (points[j] - points[m]) % P
        s = evals[j] * E.inv(den) % P
        for i in range(n):
            out[i] = (out[i] + num[i] * s) % P
    return out


Z_BOARD_COMMIT_V = [
    2426, 10710, 15244, 89073, 65613, 23051, 69107, 127496, 202311, 112438, 19493, 34450, 15808, 13514, 13555, 54715,
    147555, 27760, 6535, 13351, 1502, 4861, 6257, 33499, 3375, 29640, 70769, 59695, 1351, 28980, 86186, 13498,
    129161, 80395, 204119, 88403, 30893, 11381, 87882, 3557, 31652, 17429, 97944, 17102, 65688, 12415, 26173, 38162,
    8612, 115550, 50442, 35690, 35652, 60002, 83796, 31455, 21802, 34815, 104282, 7629, 1875, 68516, 51123, 127219,
    37335, 256888, 11747, 55662, 10852, 11424, 166053, 88854, 153304, 51360, 8139, 83706, 4929, 114588, 57071, 67335,
    15142, 12042, 178590, 1171, 24615,
]
Z_BOARD_COMMIT_R = [
    187915, 472762, 48376, 70208, 99951, 713, 34395, 12431, 40347, 10244, 87515, 28386, 1978, 316, 101853, 5228,
    31763, 64157, 49904, 68627, 90546, 24800, 17237, 52995, 3262, 17585, 14674, 74449, 75012, 104935, 6928, 89449,
    1170, 9638, 27003, 29285, 4684, 2709, 21687, 271268, 46339, 46175, 17036, 24429, 66842, 41486, 40177, 174551,
    92960, 137337, 23195, 96018, 141013, 54688, 6537, 33256, 10275, 23338, 38765, 59988, 54362, 12755, 30317, 138192,
    50707, 12098, 4942, 30676, 5252, 378500, 70207, 48665, 69166, 29218, 127059, 186479, 34813, 44267, 94673, 7323,
    130049, 127305, 42437, 16053, 74236,
]


def is_square(a):
    a %= P
    return a == 0 or pow(a, (P - 1) // 2, P) == 1


class FixedBase:
    """Window table  T[w][k] = [(k + 2) 8^w] B  (w < 84),  T[84][k] = [k 8^84 - sum_{j<84} 2 * 8^j] B,  with the Lagrange
    coefficients of k -> x(T[w][k]), the z of every window (smallest z with y + z square and -y + z non-square for all 8
    points) and u[w][k] = sqrt(y(T[w][k]) + z[w])."""

    def __init__(self, generator, z):
        assert E.is_on_curve(generator) and len(z) == NUM_WINDOWS
        self.generator = generator
        table = []
        bw, offset = generator, None          # [8^w] B, [sum_{j<w} 2 * 8^j] B
        for w in range(NUM_WINDOWS - 1):
            two = E.add(bw, bw)
            row, cur = [two], two
            for _ in range(H - 1):
                cur = E.add(cur, bw)
                row.append(cur)
            table.append(row)
            offset = E.add(offset, two)
            bw = E.add(E.add(two, two), E.add(two, two))          # 8 [8^w] B
        minus_offset = E.neg(offset)
        row, cur = [], minus_offset
        for k in range(H):
            row.append(cur)
            cur = E.add(cur, bw)
        table.append(row)
        self.table = table
        points = list(range(H))
        self.lagrange_coeffs = [lagrange_interpolate(points, [pt[0] for pt in row]) for row in table]
        self.z, self.u = list(z), []
        for row, zw in zip(table, self.z):            # find_zs_and_us' predicate, then u = sqrt(y + z)
            assert all(is_square(pt[1] + zw) and not is_square(zw - pt[1]) for pt in row), "inadmissible z"
            self.u.append([E.sqrt((pt[1] + zw) % P) for pt in row])

    def mul(self, scalar):
        return E.mul(self.generator, scalar)


@lru_cache(maxsize=None)
def board_commit_v():
    return FixedBase(BOARD_COMMIT_V, Z_BOARD_COMMIT_V)


@lru_cache(maxsize=None)
def board_commit_r():
    return FixedBase(BOARD_COMMIT_R, Z_BOARD_COMMIT_R)


def pedersen_commit(message, trapdoor):
    """R:src/utils/pedersen.rs:18-27: [message] V + [trapdoor] R with the base-field message reinterpreted as a scalar."""
    return E.add(E.mul(BOARD_COMMIT_V, message % Q), E.mul(BOARD_COMMIT_R, trapdoor % Q))
