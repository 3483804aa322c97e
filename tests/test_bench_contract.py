"""bench.py's driver contract on the CPU: the reference arm (`--impl reference`: the restated CPU prover of oracle/) prints ONE
JSON line with the keys the driver reads, and both arms build `config` from the same function (identical dicts)."""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_line(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", *extra],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    d = _reference_line()
    assert d["impl"] == "reference" and d["metric"] == "proofs_per_sec" and d["unit"] == "proofs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "restated halo2_proofs 0.2.0" in cb["sample"]
    assert 0.5 < d["cpu_native"]["native_share"] <= 1.0            # the C arithmetic, not the Python driver, is what is timed
    assert d["one_thread"]["value"] > 0 and d["one_thread"]["value"] <= d["value"] * 1.5
    assert d["config"]["workload"].startswith("batched Shot proofs (k=11, IPA/Pasta)")


def test_both_arms_share_the_config_dict():
    sys.path.insert(0, ROOT)
    import bench
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        args = bench.parse()
    finally:
        sys.argv = argv
    for name in ("shot", "board", "msm", "ntt"):
        args.workload = name
        a, b = bench.WORKLOADS[name](args), bench.WORKLOADS[name](args)
        assert bench.config_for(a) == bench.config_for(b) and "workload" in bench.config_for(a)
