"""CPU: the fixed-point view pixel -> world cell map of the observation kernel (2^-32 cell units, exact fp64 fallback inside a
guard band of 2^-19 cell) against the reference's operation sequence (agent.cpp:388-393,
grid_map.cpp:40-55) for every pixel of the view raster over random poses: outside the band the two always agree
(tests/host/fixedpoint_host_harness.cpp; 3 x 10^9 pixels hunted clean with 20 000 poses per seed)."""
import os
import subprocess

from helpers import ROOT


def test_fixed_point_cell_index_equals_exact_outside_the_guard_band(tmp_path):
    exe = tmp_path / "fx_host"
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-I", os.path.join(ROOT, "img_env_b200", "csrc"), "-o", str(exe),
                    os.path.join(ROOT, "tests", "host", "fixedpoint_host_harness.cpp")], check=True)
    r = subprocess.run([str(exe), "5", "1500"], capture_output=True, text=True)      # ~10^8 pixels, ~1.5 s
    bad, n, band = (int(x) for x in r.stdout.split())
    assert r.returncode == 0 and bad == 0 and n > 100_000_000, r.stdout
    assert 0 < band < n // 50_000, "the exact fallback must stay rare (%d of %d pixels)" % (band, n)
