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
    assert f["both"] < f["cse"] < f["plain"] and f["both"] < 0.6 * f["naive"]
