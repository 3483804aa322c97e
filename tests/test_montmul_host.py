"""CPU: the PTX carry-chain Montgomery multiplication (csrc/montmul.cuh) executed through its host emulation must
equal a*b/2^256 mod m (up to the final conditional subtraction) on random and edge operands, for both fields."""
import os, random, subprocess
from oracle import pasta

HERE = os.path.dirname(os.path.abspath(__file__))


def test_montmul_instruction_sequence_on_host(tmp_path):
    exe = str(tmp_path / "montmul_host_test")
    subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-o", exe, os.path.join(HERE, "host", "montmul_host_test.cc")], check=True)
    rnd = random.Random(2024)
    cases = []
    for f, F in ((0, pasta.FP), (1, pasta.FQ)):
        p = F.p
        edge = [0, 1, 2, p - 1, p - 2, (1 << 254), (1 << 254) - 1, F.R % p, 0xFFFFFFFF, 1 << 32, (1 << 224) - 1, (p - 1) // 2]
        for a in edge:
            for b in edge:
                cases.append((f, a, b))
        for _ in range(3000):
            cases.append((f, rnd.randrange(p), rnd.randrange(p)))
        for _ in range(200):                         # operands only need a + m < 2^256 (e.g. unreduced sums < 2^255)
            cases.append((f, rnd.randrange(1 << 255), rnd.randrange(p)))
            cases.append((f, (1 << 255) - 1 - rnd.randrange(1 << 40), p - 1 - rnd.randrange(1 << 40)))
    inp = "".join(f"{f} {a:064x} {b:064x}\n" for f, a, b in cases)
    out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split()
    assert len(out) == len(cases)
    for (f, a, b), h in zip(cases, out):
        p = (pasta.P, pasta.Q)[f]
        r = int(h, 16)
        assert r < 2 * p, "result must stay below 2m"
        assert r % p == a * b * pow(1 << 256, -1, p) % p, (f, hex(a), hex(b))
