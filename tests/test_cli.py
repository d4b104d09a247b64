"""cuopt_cli (cuopt_b200/csrc/cuopt_cli.cpp): the reference's command-line conventions (cpp/cuopt_cli.cpp: positional MPS
file, every solver parameter as --name-with-hyphens value, --relaxation; run_pdlp.cu: --path, --solution-path, mode names)
plus the replica batch mode over several files."""
import os
import subprocess

import pytest

from conftest import mps_path
from cuopt_b200 import build as b

AFIRO = "linear_programming/afiro_original.mps"


def cli(*args, timeout=300):
    exe = b.build_cli()
    return subprocess.run([exe, *args], capture_output=True, text=True, timeout=timeout)


def test_version_usage_and_host_side_errors():
    r = cli("--version")
    assert r.returncode == 0 and "cuopt-b200" in r.stdout
    assert cli().returncode == 1                                  # no file
    r = cli("/nonexistent/file.mps")
    assert r.returncode == 1 and "Parsing MPS failed" in r.stderr  # the reference's message
    r = cli(mps_path(AFIRO), "--no-such-parameter", "3")
    assert r.returncode == 1 and "unknown parameter" in r.stderr
    r = cli(mps_path(AFIRO), "--time-limit")
    assert r.returncode == 1 and "needs a value" in r.stderr


@pytest.mark.gpu
def test_solves_afiro_with_reference_style_options(tmp_path):
    sol = tmp_path / "afiro.sol"
    r = cli(mps_path(AFIRO), "--method", "1", "--optimality-tolerance", "1e-8", "--pdlp-solver-mode", "Stable2",
            "--solution-path", str(sol))
    assert r.returncode == 0, r.stderr
    assert "Status: Optimal" in r.stdout and "-4.647531" in r.stdout  # -464.7531428 to the PDLP tolerance asked for
    assert sol.exists() and "X01" in sol.read_text()


@pytest.mark.gpu
def test_batch_mode_and_relaxation():
    files = [mps_path(AFIRO), mps_path("mip/sudoku.mps"), mps_path("mip/sample.mps")]
    r = cli(*files, "--relaxation", "--method", "1", "--gpus", "2")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if "Status:" in ln]
    assert len(lines) == 3 and all("Optimal" in ln for ln in lines)
    # without --relaxation the integer problems are refused (LP-only build), afiro still solves, exit code 1
    r = cli(*files, "--method", "1")
    assert r.returncode == 1 and "integer variables" in r.stderr
