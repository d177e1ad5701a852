"""CPU: world2cell_fast (img_env_b200/csrc/tfmath.cuh) -- the division-free cell index of the footprint / collision kernels --
equals the reference's int(round(x / res)) (GridMap::world2map, grid_map.cpp:40-44) on 10 M coordinates, including every
rounding boundary of an 8192-cell map approached to within a few ulp from both sides (tests/host/tfmath_host_harness.cpp)."""
import os
import subprocess

from helpers import ROOT


def test_world2cell_fast_equals_round_of_quotient(tmp_path):
    exe = tmp_path / "tf_host"
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-I", os.path.join(ROOT, "img_env_b200", "csrc"), "-o", str(exe),
                    os.path.join(ROOT, "tests", "host", "tfmath_host_harness.cpp")], check=True)
    for seed in (1, 2, 3):
        r = subprocess.run([str(exe), str(seed)], capture_output=True, text=True)
        bad, n = (int(x) for x in r.stdout.split())
        assert r.returncode == 0 and bad == 0 and n > 10_000_000, r.stdout
