"""CPU: the gate-DAG compiler of the h(X) evaluator (csrc/evalprog.h: common sub-expressions once, common factors
hoisted) emits programs that evaluate to sum_e y^(E-1-e) expr_e on random expression families, in every combination of
the two rewrites, within the evaluator's stack / temporary limits; and the rewrites do reduce the multiplication count."""
import os, subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_gate_dag_compiler_on_host(tmp_path):
    exe = str(tmp_path / "evalprog_host_test")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(HERE, "host", "evalprog_host_test.cc")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == "ok", out
    f = dict(zip(out[1::2], map(int, out[2::2])))
    assert f["both"] < f["cse"] < f["plain"] and f["both"] < 0.6 * f["naive"] and f["nested"] < f["cse"] and f["canon"] < f["cse"]


def _dump_gates(cs, path):
    """Gate polynomials of a ConstraintSystem mirror as token lines (same flattening and tier rule as the library:
    plonk/prover.py `flatten_circuit`, csrc/prover.cu `build_programs`)."""
    from battlezips_halo2_b200.plonk.prover import flatten_circuit
    ir = cs.to_ir()
    toks, consts = [], {}

    def cidx(v):
        return consts.setdefault(v % ir["modulus"], len(consts))

    def emit(e, out):
        kind = e[0]
        if kind == "const": out.append((0, cidx(e[1]), 0))
        elif kind in ("advice", "fixed", "instance"): out.append(({"advice": 1, "fixed": 2, "instance": 3}[kind], e[1], e[2]))
        elif kind == "neg": emit(e[1], out); out.append((4, 0, 0))
        elif kind == "sum": emit(e[1], out); emit(e[2], out); out.append((5, 0, 0))
        elif kind == "product": emit(e[1], out); emit(e[2], out); out.append((6, 0, 0))
        elif kind == "scaled": emit(e[1], out); out.append((7, cidx(e[2]), 0))

    def degree(e):
        kind = e[0]
        if kind == "const": return 0
        if kind in ("advice", "fixed", "instance"): return 1
        if kind in ("neg", "scaled"): return degree(e[1])
        if kind == "sum": return max(degree(e[1]), degree(e[2]))
        return degree(e[1]) + degree(e[2])

    ext = 1
    while (1 << ext) < cs.degree() - 1:
        ext += 1
    polys = []
    for gate in ir["gates"]:
        for poly in gate["polys"]:
            out = []
            emit(poly, out)
            d, t = max(1, degree(poly)), 0
            while t + 1 < 3 and ((1 << ext) >> (t + 1)) >= 1 and (d - 1) <= ((1 << ext) >> (t + 1)):
                t += 1
            polys.append((t, out))
    n_exprs = len(polys) + 64                      # the argument terms follow the gates; only the count matters here
    with open(path, "w") as f:
        f.write(f"{n_exprs} {cs.num_advice} {len(consts)}\n")
        for e, (t, out) in enumerate(polys):
            f.write(f"{t} {e} {len(out)} " + " ".join(f"{op} {a} {b}" for op, a, b in out) + "\n")
    return len(polys)


def test_gate_dag_compiler_on_the_shot_and_board_constraint_systems(tmp_path):
    """The real gate sets: the compiled programs are correct, fit the evaluator's stack (EVAL_STACK) and temporaries
    (EVAL_TMP), and the DAG saves multiplications in the tier that is evaluated on every point of the extended coset."""
    from battlezips_halo2_b200.circuits import shot_circuit, board_circuit
    exe = str(tmp_path / "evalprog_host_test")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(HERE, "host", "evalprog_host_test.cc")], check=True)
    for name, make, npolys in (("shot", shot_circuit, 82), ("board", board_circuit, 141)):
        cs = make(0)[0]                              # after selector compression: what keygen hands the device
        path = str(tmp_path / f"{name}.gates")
        assert _dump_gates(cs, path) == npolys
        out = subprocess.run([exe, path], capture_output=True, text=True).stdout.strip().splitlines()
        assert out[-1] == "ok", out
        rows = [dict(zip(l.split()[0::2], map(int, l.split()[1::2]))) for l in out[:-1]]
        print(name, rows)
        assert rows[0]["dag_muls"] < 0.4 * rows[0]["tree_muls"] and rows[0]["derived"] >= 10          # compressed selectors leave the per-point program
        assert sum(r["dag_muls"] for r in rows) < sum(r["tree_muls"] for r in rows)


def test_generated_quotient_kernels_are_current():
    """csrc/gen_quotient.cu (h(X) of Shot / Board as straight-line code) is what scripts/gen_quotient_kernels.py emits for the
    current circuits and compiler: a stale file would silently fall back to the interpreter (the hash lookup misses)."""
    import sys
    root = os.path.dirname(HERE)
    assert subprocess.run([sys.executable, os.path.join(root, "scripts", "gen_quotient_kernels.py"), "--check"]).returncode == 0, \
        "csrc/gen_quotient.cu is stale: run python scripts/gen_quotient_kernels.py and rebuild"
