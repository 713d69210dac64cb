"""Host emulation of device arithmetic (tests/host_emul, g++).  The two-wide main-pass lane EarLane2
(the experimental f32x2 kernels, NELE_F32X2=1) must agree with two scalar EarLane<float> lanes.  The
device build runs the same source with fma.rn.f32x2 in place of the component-wise fmaf."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_two_wide_lane_matches_the_scalar_lanes(tmp_path):
    exe = str(tmp_path / "ear_x2_emul")
    src = os.path.join(ROOT, "tests", "host_emul", "ear_x2_emul.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe, src], check=True)
    out = subprocess.run([exe, "18000"], check=True, stdout=subprocess.PIPE, text=True).stdout
    worst = float(re.search(r"WORST ([0-9.eE+-]+)", out).group(1))
    levels = [float(m) for m in re.findall(r"max level ([0-9.]+) dB", out)]
    assert len(levels) == 11 and min(levels) > 20.0           # the lanes were driven, not silent
    # the scalar lanes round a * r + x twice with contraction off, the two-wide lane once (fmaf): the
    # recurrences differ by rounding only -- far below the model's own 0.1 dB dither
    assert worst < 2e-3, out


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_in_place_fft400_matches_a_direct_dft(tmp_path):
    """fft400.cuh (SIIB's 400-point STFT, intel.py:52-54): both phases run in place on one buffer and leave
    X[k] at fft400_pos(k).  Against a float64 DFT, and independent of the order the lanes run in."""
    exe = str(tmp_path / "fft400_emul")
    src = os.path.join(ROOT, "tests", "host_emul", "fft400_emul.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src], check=True)
    out = subprocess.run([exe], check=True, stdout=subprocess.PIPE, text=True).stdout
    err = float(re.search(r"MAXERR ([0-9.eE+-]+)", out).group(1))
    order = float(re.search(r"ORDER ([0-9.eE+-]+)", out).group(1))
    assert err < 1e-6, out          # float32 butterflies: about 1.5e-7 of the largest bin
    assert order == 0.0, out        # no lane reads what another lane of the same phase writes


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_sturm_count_and_bisection_plan_match_the_ratio_form(tmp_path):
    """sturm.cuh (the eigenvalue bisection of siib_trieig_kernel / siib_smallvec_kernel, pysiib's KLT): sign masks,
    16-step blocks over the padded table and the exponent-field range guard give the same counts as the ratio-form
    Sturm count in long double -- on dense, graded (14 decades), numerically rank-deficient, Toeplitz, decoupled
    (underflow side) and growing (overflow side) tridiagonals of both kernel sizes -- and the grid start + k steps
    land on the eigenvalues of plain bisection."""
    exe = str(tmp_path / "sturm_emul")
    src = os.path.join(ROOT, "tests", "host_emul", "sturm_emul.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src], check=True)
    out = subprocess.run([exe], check=True, stdout=subprocess.PIPE, text=True).stdout
    rows = re.findall(r"(\S+) n=(\d+) points=(\d+) skipped=(\d+) MISMATCH=(\d+) EIGERR=([0-9.eE+-]+)", out)
    assert len(rows) == 12, out
    for name, n, points, skipped, mismatch, eigerr in rows:
        assert int(points) > 700 and int(skipped) < 10, out
        assert int(mismatch) == 0, out
        # half the final interval: 6 / 420 * 2^-38 / 2 = 2.6e-14 (n = 420), 6 / 112 * 2^-52 / 2 = 6e-18 (+ rounding of mid)
        assert float(eigerr) < (3e-14 if n == "420" else 1e-15), out
