"""CPU-side checks of bench.py: the term count behind `value`, the reference arm's JSON contract (run here on a small
lmax), and that the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _bench(*args, timeout=300):
    env = dict(os.environ, OMP_NUM_THREADS="1")            # what torchrun exports; the CPU arm must override it
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT, env=env)


def test_term_count_matches_the_survey_table(ps):
    """T_fam(lmax) = sum_l1 (2 l1 + 1)(lmax - l1 + 1): SURVEY.md 8d table; the library counts the same (psb200_terms)."""
    sys.path.insert(0, ROOT)
    import bench
    for lmax, expect in ((767, 1.5129e8), (2508, 5.2679e9), (3071, 9.6684e9), (6143, 7.7328e10), (12287, 6.1855e11)):
        t = bench.t_fam(lmax)
        assert abs(t / expect - 1) < 1e-4
        assert ps.lib().psb200_terms(1, lmax, 0, lmax + 1) == t
    assert bench.t_fam(6143, 0, 100) + bench.t_fam(6143, 100, 6144) == bench.t_fam(6143)
    assert sum(j[3] for j in bench.JOBS) == 7              # reference families per step (TT 1, EE/BB 2, TTTT 1, EEEE 1, TETE 2)


def test_reference_arm_contract():
    out = _bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--lmax", "767")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "terms/s" and line["higher_is_better"] is True
    assert line["dtype"] == "f64" and line["vs_baseline"] is None and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and "sample" in cb
    assert cb["cores"] == (os.cpu_count() or 1) or cb["cores"] >= 1      # every host thread, despite OMP_NUM_THREADS=1
    assert cb["cores"] > 1 or (os.cpu_count() or 1) == 1
    e = line["e2e"]
    assert e["value"] == line["value"] and e["unit"] == line["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert line["value"] > 1e7


def test_product_arm_needs_a_device(ps):
    if ps.lib().psb200_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    out = _bench("--steps", "1", "--warmup", "0", "--lmax", "255", "--no-cpu", "--no-extra", timeout=600)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)
