"""No-GPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/imgenv.h declares, and fails loudly (never silently falls back) without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import ROOT, base_cfg, build_spec


@pytest.fixture(scope="module")
def lib():
    from img_env_b200.build import build
    from img_env_b200.lib import load_library
    build()
    return load_library()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "imgenv.h")).read()
    names = sorted(set(re.findall(r"\b(imgenv_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), "missing export " + n


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.imgenv_version()
    assert isinstance(lib.imgenv_last_error(), bytes)


def test_f16_lut_matches_numpy(lib):
    lut = np.zeros(256, np.uint16)
    lib.imgenv_f16_lut(lut.ctypes.data_as(C.c_void_p))
    want = (np.arange(256, dtype=np.uint8).astype("float16") / 255.0).view(np.uint16)   # yaml_env.py:438
    assert np.array_equal(lut, want)


def test_yaw_from_quaternion_matches_tf_model(lib):
    from img_env_b200.spec import rpy_to_q
    import math
    for yaw in np.linspace(-3.14, 3.14, 41):
        q = rpy_to_q(float(yaw))
        got = lib.imgenv_yaw_from_quaternion(*[C.c_double(v) for v in q])
        assert abs(math.remainder(got - yaw, 2 * math.pi)) < 1e-12


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from img_env_b200.lib import BatchedSim
    spec = build_spec(base_cfg())
    with pytest.raises(RuntimeError):
        BatchedSim(spec)
    # and the raw C entry point reports the reason instead of computing on the CPU
    from img_env_b200.lib import ImgenvConfig
    cfg = ImgenvConfig(num_scenes=1, num_robots=1)
    h = C.c_void_p()
    grid = np.zeros((4, 4), np.uint8); d = np.zeros(25); s = np.zeros(1)
    rc = lib.imgenv_create(C.byref(cfg), grid.ctypes.data_as(C.c_void_p), 4, 4, d.ctypes.data_as(C.c_void_p), None,
                           s.ctypes.data_as(C.c_void_p), 0, C.byref(h))
    assert rc != 0 and b"CUDA" in lib.imgenv_last_error()


def test_spec_follows_reference_request_building():
    cfg = base_cfg(R=2, P=3, scene="ervoscene")
    spec = build_spec(cfg)
    assert spec["grid"].shape == (733, 733)                      # int(110 * f32(0.1) / f32(0.015)), grid_map.cpp:31-32
    assert spec["scalars"][17] == 0.0 and spec["scalars"][18] == 0.0   # beep_r / ped_ca_p never reach the wire
    assert build_spec(cfg, opt_in_beep=True)["scalars"][17] == 1.0
    assert list(spec["ped_desc"][0][:7]) == [2, 0, 0.1, 0.1, 0, -0.1, 0.1]   # leg: right = (x, -y, r), reset_helper.py:399-403
    assert spec["robot_size_last"] == [0.17, 0.17]


def test_ctypes_structs_match_the_c_header(tmp_path):
    """sizeof / offsetof of imgenv_config and imgenv_outputs as a C compiler sees include/imgenv.h == the ctypes mirror."""
    import ctypes as C
    import subprocess
    from img_env_b200.lib import ImgenvConfig, ImgenvOutputs
    src = tmp_path / "layout.c"
    fields_cfg = [f[0] for f in ImgenvConfig._fields_]
    fields_out = [f[0] for f in ImgenvOutputs._fields_]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "imgenv.h"', 'int main(void) {',
             '  printf("%zu %zu\\n", sizeof(imgenv_config), sizeof(imgenv_outputs));']
    lines += ['  printf("%%zu\\n", offsetof(imgenv_config, %s));' % f for f in fields_cfg]
    lines += ['  printf("%%zu\\n", offsetof(imgenv_outputs, %s));' % f for f in fields_out]
    lines += ['  return 0; }']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert [int(got[0]), int(got[1])] == [C.sizeof(ImgenvConfig), C.sizeof(ImgenvOutputs)]
    want = [getattr(ImgenvConfig, f).offset for f in fields_cfg] + [getattr(ImgenvOutputs, f).offset for f in fields_out]
    assert [int(x) for x in got[2:]] == want


def test_header_is_plain_c_and_links(tmp_path, lib):
    """include/imgenv.h compiles as strict C99 and a C program links against the shared library (no torch, no C++)."""
    import subprocess
    from img_env_b200.lib import LIB_PATH
    src = tmp_path / "client.c"
    src.write_text('#include <stdio.h>\n#include "imgenv.h"\n'
                   'int main(void) { imgenv_t* h = 0; imgenv_config cfg = {0};\n'
                   '  printf("%s\\n", imgenv_version());\n'
                   '  int rc = imgenv_create(&cfg, 0, 0, 0, 0, 0, 0, 0, &h);   /* null arguments: must fail cleanly */\n'
                   '  printf("%d %s\\n", rc, imgenv_last_error());\n'
                   '  return rc < 0 ? 0 : 1; }\n')
    exe = tmp_path / "client"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src),
                    LIB_PATH, "-Wl,-rpath," + os.path.dirname(LIB_PATH)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "img_env_b200" in r.stdout and "null argument" in r.stdout
