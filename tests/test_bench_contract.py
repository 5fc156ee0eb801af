"""bench.py's reference arm runs without a GPU (it times the unmodified reference from baseline/_ref on the host cores,
or the oracle port where no reference tree exists): check the one-JSON-line contract the driver parses — exactly one
line on stdout, the required keys, the tier's cpu_baseline / e2e objects, and the fallback when baseline/_ref is absent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_reference_arm(env=None):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                        "--warmup", "1", "--cpu-sample-img", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["unit"] == "bodies/s" and "workload" in d["config"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    return d


def test_reference_arm_prints_one_contract_line():
    """With a reference tree available (baseline/_ref, made by tools/install_ref.sh / build()) the arm drives the unmodified
    reference; its `config` equals the GPU arm's (bench.CONFIG), so the driver's same_config check holds."""
    sys.path.insert(0, ROOT)
    import bench
    from baseline import ref_harness
    d = _run_reference_arm()
    assert d["config"] == json.loads(json.dumps(bench.CONFIG))
    expect = "reference" if ref_harness.reference_root()[0] else "port"
    assert d["cpu_baseline"]["kind"] == expect, d["cpu_baseline"]


def test_reference_arm_falls_back_to_the_oracle_port_without_a_reference_tree(tmp_path):
    env = dict(os.environ, EHB_REFERENCE_ROOT=str(tmp_path), EHB_IGNORE_REFERENCE="1")
    d = _run_reference_arm(env)
    assert d["cpu_baseline"]["kind"] == "port"


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""
