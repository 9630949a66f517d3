"""Fuzzes the claim the sub-tile culling of the render kernels rests on -- block_candidate() == false implies that no ray of the block is
accepted by the exact test -- by compiling the kernels' own arithmetic (3dgrut_b200/csrc/subtile_cull.cuh, hit_math.cuh) with g++
(tests/host_emul/cull_host.cpp).  Random 8x4 ray blocks (pitch 1e-4 .. 3e-2 rad, normalised or not) x random Gaussians (scales 1e-4 .. 3
with anisotropy up to 30:1, depth 0.05 .. 50, densities down to the alpha threshold).  This test found a real bug in round 1: for
needle-like Gaussians the determinant of the quadratic overflowed and true hits were culled (fixed by normalising the coefficients)."""
import ctypes as C
import os
import subprocess

import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")
ROOT = os.path.dirname(os.path.dirname(HERE))


def _lib():
    so, src = os.path.join(HERE, "libcull_host.so"), os.path.join(HERE, "cull_host.cpp")
    deps = [src] + [os.path.join(ROOT, "3dgrut_b200", "csrc", f) for f in ("subtile_cull.cuh", "hit_math.cuh", "gut_common.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        cxx = "/opt/gcc/bin/g++" if os.path.exists("/opt/gcc/bin/g++") else "g++"
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
            pytest.skip("CUDA headers not found")
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-w", f"-I{cuda_inc}", "-shared", "-o", so, src])
    lib = C.CDLL(so)
    lib.cull_fuzz.restype = C.c_int64
    return lib


@pytest.mark.parametrize("degree", [2, 4])
@pytest.mark.parametrize("slack", [1.0, 0.99998])  # 0.99998: 20x the relative error of the approximate exp / division on the GPU
def test_culling_never_drops_a_ray_the_exact_test_accepts(degree, slack):
    lib = _lib()
    stats, worst = (C.c_int64 * 3)(), (C.c_float * 12)()
    cases = 1_500_000
    bad = lib.cull_fuzz(C.c_uint64(7 + degree), C.c_int64(cases), C.c_int(degree), C.c_float(0.0113), C.c_float(1 / 255), C.c_float(0.99),
                        C.c_float(slack), stats, worst)
    print(f"[cull fuzz] degree {degree} slack {slack}: {cases} cases, {stats[0]} with an accepted ray, {stats[1]} culled, "
          f"{stats[2]} kept without a hit, {bad} violations")
    assert bad == 0, f"first violating particle record: {list(worst)}"
    assert stats[0] > 100_000 and stats[1] > 500_000  # the fuzz really exercises both outcomes
