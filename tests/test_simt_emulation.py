"""Warp-cooperative per-tet insertion (csrc/ia_complex_warp.cuh, csrc/iso_record.cuh) run on the CPU.

32 host threads play the lanes of one warp (tests/simt/simt_emul.h); the result of every insertion and the iso
record are compared with the serial device code compiled for the host, which the GPU parity tests pin to the
oracle.  A ThreadSanitizer build checks that every cross-lane dependency is covered by a __syncwarp()
(ballots / shuffles are emulated with relaxed atomics, so they do not hide a missing barrier).
"""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "simt", "ia_warp_check.cpp")
SRC_MI = os.path.join(HERE, "simt", "mi_warp_check.cpp")
FLAGS = ["-std=c++17", "-O1", "-g", "-ffp-contract=off", "-pthread"]


def _build(tmp_path, extra, name, src=SRC):
    exe = str(tmp_path / name)
    r = subprocess.run(["g++", *FLAGS, *extra, "-o", exe, src], capture_output=True, text=True, cwd=os.path.dirname(src))
    return exe, r


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_warp_insertion_equals_serial(tmp_path):
    exe, r = _build(tmp_path, [], "ia_warp_check")
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([exe, "400", "21"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert out.stdout.count("mismatches 0") == 3, out.stdout


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_warp_insertion_is_race_free(tmp_path):
    exe, r = _build(tmp_path, ["-fsanitize=thread"], "ia_warp_check_tsan")
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-300:])
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0")
    out = subprocess.run([exe, "80", "5"], capture_output=True, text=True, timeout=900, env=env)
    assert "ThreadSanitizer" not in out.stderr, out.stderr[:3000]
    assert out.returncode == 0, out.stdout[-2000:]


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_warp_material_insertion_equals_serial(tmp_path):
    """csrc/mi_complex_warp.cuh vs MIComplex::add_material: ties, duplicates, near-duplicates, whole-face ties."""
    exe, r = _build(tmp_path, [], "mi_warp_check", SRC_MI)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([exe, "400", "33"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert out.stdout.count("mismatches 0") == 2, out.stdout


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_warp_material_insertion_is_race_free(tmp_path):
    exe, r = _build(tmp_path, ["-fsanitize=thread"], "mi_warp_check_tsan", SRC_MI)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-300:])
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0")
    out = subprocess.run([exe, "80", "6"], capture_output=True, text=True, timeout=900, env=env)
    assert "ThreadSanitizer" not in out.stderr, out.stderr[:3000]
    assert out.returncode == 0, out.stdout[-2000:]
