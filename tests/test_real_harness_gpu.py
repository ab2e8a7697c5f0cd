"""The UNMODIFIED NPBench CLI with `-f b200` (BASELINE.json configs[0]).

`python -m npbench_b200.run -- -b <bench> -f b200 -p <preset> -r 3` executes the reference's own
run_benchmark.py (run_benchmark.py:49-57) -> Test.run (npbench/infrastructure/test.py:53-163): NumPy
validation run, first b200 execution, util.validate (utilities.py:154-180), the timed repetitions
(timeit, utilities.py:135-151) and the sqlite rows.  The checkout is /root/reference in the build
container and the byte-for-byte staged copy baseline/_ref on the GPU box (npbench_b200/overlay.py).

CPU part: the staged copy is unmodified and the harness runs from it with `-f numpy`.
GPU part: every stencil kernel x preset validates against NumPy inside the harness and leaves
`framework='b200', validated=1` rows in npbench.db.
"""
import os
import sqlite3
import subprocess
import sys

import pytest

from conftest import ROOT
from npbench_b200 import overlay


def _ref_or_skip():
    try:
        return overlay.find_reference()
    except FileNotFoundError as e:
        pytest.skip(str(e))


def run_cli(tmp_path, ref, *cli, timeout=900):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "npbench_b200.run", "--reference", ref, "--overlay",
                        str(tmp_path / "ov"), "--"] + list(cli),
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=timeout)
    return r


def rows(tmp_path, framework):
    con = sqlite3.connect(str(tmp_path / "npbench.db"))
    try:
        return con.execute("SELECT benchmark, preset, framework, version, details, validated, time FROM results "
                           "WHERE framework = ?", (framework,)).fetchall()
    finally:
        con.close()


@pytest.mark.skipif(not overlay.is_checkout("/root/reference"), reason="build container only")
def test_staged_copy_is_byte_identical_to_the_reference(tmp_path):
    dest = overlay.stage_reference("/root/reference", str(tmp_path / "_ref"))
    assert overlay.verify_staged(dest) == []
    for rel in ("npbench/infrastructure/test.py", "npbench/infrastructure/framework.py", "run_benchmark.py",
                "npbench/benchmarks/polybench/jacobi_2d/jacobi_2d_numpy.py", "bench_info/heat_3d.json"):
        assert open(os.path.join(dest, rel), "rb").read() == open(os.path.join("/root/reference", rel), "rb").read()
    # no plugin file leaks into the staged copy: it is the reference, not the overlay
    assert not os.path.exists(os.path.join(dest, "framework_info", "b200.json"))


def test_harness_runs_numpy_from_the_staged_copy(tmp_path):
    if not overlay.is_checkout(overlay.STAGED):
        pytest.skip("baseline/_ref not staged (run __graft_entry__.build() where /root/reference is mounted)")
    assert overlay.verify_staged(overlay.STAGED) == []
    r = run_cli(tmp_path, overlay.STAGED, "-b", "jacobi_2d", "-f", "numpy", "-p", "S", "-r", "2")
    assert r.returncode == 0 and "NumPy - default - median" in r.stdout, r.stdout + r.stderr
    got = rows(tmp_path, "numpy")
    assert len(got) == 2 and all(g[0] == "jacobi2d" and g[1] == "S" for g in got)


STENCILS = ["jacobi_2d", "heat_3d", "fdtd_2d", "hdiff", "vadv"]
SHORT = {"jacobi_2d": "jacobi2d", "heat_3d": "heat3d", "fdtd_2d": "fdtd_2d", "hdiff": "hdiff", "vadv": "vadv",
         "jacobi_1d": "jacobi1d", "seidel_2d": "seidel2d", "adi": "adi", "cavity_flow": "cavflow",
         "channel_flow": "chanflow"}


def _check_b200_run(tmp_path, bench, preset, repeat):
    ref = _ref_or_skip()
    r = run_cli(tmp_path, ref, "-b", bench, "-f", "b200", "-p", preset, "-r", str(repeat))
    out = r.stdout + r.stderr
    assert r.returncode == 0, out
    assert "validation: SUCCESS" in r.stdout, out            # test.py:118
    assert "did not validate" not in out and "Failed to" not in out, out
    got = rows(tmp_path, "b200")
    assert len(got) == repeat, (got, out)
    for benchmark, pre, fw, version, details, validated, t in got:
        assert pre == preset and fw == "b200" and validated == 1 and t > 0.0
        assert version.startswith("0.") and "lib" in version
    # the NumPy validation run of the same invocation is not stored (test.py:66-68), only b200 rows
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["S", "M", "L"])
@pytest.mark.parametrize("bench", STENCILS)
def test_real_harness_b200(bench, preset, tmp_path):
    _check_b200_run(tmp_path, bench, preset, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("bench", ["hdiff", "vadv", "fdtd_2d"])
def test_real_harness_b200_paper(bench, tmp_path):
    # NumPy's validation run at `paper`: hdiff 0.5 s, vadv 2 s, fdtd_2d 7 s (jacobi_2d 167 s and
    # heat_3d 35 s are covered against the oracle in test_parity_gpu.py instead)
    _check_b200_run(tmp_path, bench, "paper", 2)


@pytest.mark.gpu
@pytest.mark.parametrize("bench", ["jacobi_1d", "seidel_2d", "adi", "cavity_flow", "channel_flow"])
def test_real_harness_b200_widening_row(bench, tmp_path):
    _check_b200_run(tmp_path, bench, "S", 2)


@pytest.mark.gpu
@pytest.mark.parametrize("bench,preset", [("hdiff", "S"), ("hdiff", "paper"), ("vadv", "M")])
def test_real_harness_b200_column_shards(bench, preset, tmp_path, monkeypatch):
    """BASELINE.json configs[2]: hdiff `paper` column-sharded, reached through the UNMODIFIED harness with
    NPB_B200_GPUS=N (one process, N device slots; on a one-GPU box the slots share device 0)."""
    monkeypatch.setenv("NPB_B200_GPUS", "3")
    _check_b200_run(tmp_path, bench, preset, 2)
