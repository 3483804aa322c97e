"""Extract the reference's own Pallas golden vectors into a committed fixture.

Source (read-only, this container only):
  /root/reference/src/utils/constants/fixed_bases/board_commit_v.rs:5-14 (GENERATOR),
  :17-26 (Z), :28-2919 (U); same lines of board_commit_r.rs.
  Hash-to-curve inputs: /root/reference/src/utils/constants.rs:6,12,15.
Run:  python tests/golden/make_pallas_kats.py   (writes pallas_fixed_base_kats.json)
Only literal DATA is extracted (byte arrays / integers); no reference code is copied.
"""
import json, re, os

REF = "/root/reference/src/utils/constants/fixed_bases"
OUT = os.path.join(os.path.dirname(__file__), "pallas_fixed_base_kats.json")


def ints(s):
    return [int(x) for x in re.findall(r"\d+", s)]


def parse(path):
    src = open(path).read()
    gen = re.search(r"pub const GENERATOR[^=]*=\s*\((.*?)\);", src, re.S).group(1)
    g = ints(gen)
    assert len(g) == 64
    z = ints(re.search(r"pub const Z: \[u64; NUM_WINDOWS\] = \[(.*?)\];", src, re.S).group(1))
    assert len(z) == 85
    u_src = re.search(r"pub const U: \[\[\[u8; 32\]; H\]; NUM_WINDOWS\] = \[(.*?)\n\];", src, re.S).group(1)
    u = ints(u_src)
    assert len(u) == 85 * 8 * 32, len(u)
    return {
        "generator_x": bytes(g[:32]).hex(),
        "generator_y": bytes(g[32:]).hex(),
        "z": z,
        "u": [[bytes(u[(w * 8 + k) * 32:(w * 8 + k + 1) * 32]).hex() for k in range(8)] for w in range(85)],
    }


if __name__ == "__main__":
    out = {
        "personalization": "battlezips:hash2curve",
        "v": dict(message="v", **parse(f"{REF}/board_commit_v.rs")),
        "r": dict(message="r", **parse(f"{REF}/board_commit_r.rs")),
    }
    json.dump(out, open(OUT, "w"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
