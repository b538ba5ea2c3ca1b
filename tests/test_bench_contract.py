"""bench.py contract (CPU side): the reference arm prints exactly one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("kind", ["port", "auto"])
def test_reference_arm_line(oracle, kind):
    """--impl reference: the unmodified Numba reference (live tree or the bytecode staged in oracle/_ref)
    when it is available, else the C port; either way one JSON line on the B200 arm's config."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3",
                          "--warmup", "1", "--bodies-per-gpu", "65536", "--ref-kind", kind],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-1500:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "body-force updates/sec" and d["unit"] == "bodies/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and "workload" in d["config"]
    from oracle import ref_numba
    want = "reference" if (kind == "auto" and ref_numba.available()) else "port"
    assert d["cpu_baseline"]["kind"] == want and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["bodies_per_gpu"] == 65536 and d["steps"] == 3  # the B200 arm's config, not a capped one
    assert d["e2e"] == {"value": d["value"], "unit": "bodies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e5 and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent(oracle):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode != 0 and "no CPU path" in (res.stderr + res.stdout)
    assert res.stdout.strip() == ""  # no fabricated line


def test_tools_and_harnesses_compile():
    """Every stand-alone script (tools/, tests/harness/, examples/) at least byte-compiles."""
    import glob
    import py_compile

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    scripts = [p for d in ("tools", os.path.join("tests", "harness"), "examples")
               for p in glob.glob(os.path.join(root, d, "*.py"))]
    assert len(scripts) >= 15
    for p in scripts:
        py_compile.compile(p, doraise=True, cfile=os.path.join("/tmp", "h2o_pyc_" + os.path.basename(p) + "c"))
