"""CPU: the binary-GCD inversion written for round 2 (csrc/gcdinv.h, host/device code without PTX) equals pow(a, -1, p)
on random and edge operands for both Pasta fields, and maps 0 to 0 (ff::BatchInvert's convention)."""
import os, random, subprocess
from oracle import pasta

HERE = os.path.dirname(os.path.abspath(__file__))


def test_binary_gcd_inverse_on_host(tmp_path):
    exe = str(tmp_path / "gcdinv_host_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "host", "gcdinv_host_test.cc")], check=True)
    rnd = random.Random(77)
    cases = []
    for f, p in ((0, pasta.P), (1, pasta.Q)):
        edge = [0, 1, 2, 3, p - 1, p - 2, (p + 1) // 2, (p - 1) // 2, 1 << 32, 1 << 64, 1 << 224, (1 << 254), (1 << 254) - 1,
                0xFFFFFFFF, (1 << 32) * 12345, (1 << 96) * 3, p - (1 << 32), p - (1 << 200)]
        cases += [(f, a % p) for a in edge]
        cases += [(f, (1 << s) % p) for s in range(0, 255, 7)]
        cases += [(f, rnd.randrange(1, p)) for _ in range(4000)]
        cases += [(f, rnd.randrange(1, 1 << rnd.randrange(1, 254))) for _ in range(500)]
    inp = "".join(f"{f} {a:064x}\n" for f, a in cases)
    out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split()
    assert len(out) == len(cases)
    for (f, a), h in zip(cases, out):
        p = (pasta.P, pasta.Q)[f]
        r = int(h, 16)
        assert r < p
        assert r == (pow(a, -1, p) if a else 0), (f, hex(a))
