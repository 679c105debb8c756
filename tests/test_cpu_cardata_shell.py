"""CPU parity of the host-side CarData helpers the ROS nodes call for their markers (createVehicleGeometry,
getVehicleGeometry_state, getCircles, getBoundingCircle, local2Global, getVehicleStep): the shell's output against the
reference's own CarData.cpp (compiled from the reference checkout with the oracle's message shims), printed with %.17g for
the same pseudo-random states and compared as text.  Skipped without the reference checkout."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HPP_REF", "/root/reference")

BODY = r'''
#include <cstdio>
#include <cstdint>
int main() {
  CarData car(0.6, 2.75, 1.1, 1.2, 1.9);
  car.createVehicleGeometry();
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
  for (int i = 0; i < 200; i++) {
    State st(rnd() * 80 - 40, rnd() * 80 - 40, rnd() * 12 - 6);
    auto poly = car.getVehicleGeometry_state(st);
    for (auto &p : poly.points) std::printf("P %.9g %.9g\n", (double)p.x, (double)p.y);
    for (auto &c : car.getCircles(st)) std::printf("C %.17g %.17g %.17g\n", c.x, c.y, c.r);
    auto b = car.getBoundingCircle(st);
    std::printf("B %.17g %.17g %.17g\n", b.x, b.y, b.r);
    State g = car.local2Global(st, State(rnd() * 4 - 2, rnd() * 4 - 2, rnd()));
    std::printf("G %.17g %.17g %.17g\n", g.x, g.y, g.heading);
    State n = car.getVehicleStep(st, rnd() * 1.2 - 0.6, i % 2 ? 1 : -1, 1.0);
    std::printf("S %.17g %.17g %.17g\n", n.x, n.y, n.heading);
  }
  return 0;
}
'''


@pytest.mark.parametrize("pkg", ["reeds_shepp_hybrid_astar", "dubins_hybrid_star"])
def test_cardata_shell_matches_the_reference(pkg):
    src = os.path.join(REF, pkg, "src", "CarData.cpp")
    if not os.path.exists(src):
        pytest.skip("reference checkout not present")
    has_circles = "getCircles" in open(os.path.join(REF, pkg, "include", "CarData.h")).read()
    body = BODY if has_circles else "\n".join(ln for ln in BODY.splitlines()
                                              if not any(k in ln for k in ("getCircles", "getBoundingCircle", "local2Global",
                                                                           '"B ', '"G ')))
    tmp = tempfile.mkdtemp()
    try:
        with open(os.path.join(tmp, "ref.cpp"), "w") as f:
            f.write('#include "CarData.h"\n' + body)
        with open(os.path.join(tmp, "mine.cpp"), "w") as f:
            f.write('#include "CarDataShell.h"\n' + body)
        shim = os.path.join(ROOT, "oracle", "ref_shim")
        flags = ["-std=c++17", "-O2", "-ffp-contract=off", "-w"]
        r = subprocess.run(["g++", *flags, "-include", os.path.join(shim, "ros_std_includes.hpp"), "-I" + shim,
                            "-I" + os.path.join(REF, pkg, "include"), os.path.join(tmp, "ref.cpp"), src,
                            "-o", os.path.join(tmp, "ref")], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        r = subprocess.run(["g++", *flags, "-I" + os.path.join(ROOT, "hybrid_path_planning_b200", "host"),
                            "-I" + os.path.join(ROOT, "include"), os.path.join(tmp, "mine.cpp"), "-o", os.path.join(tmp, "mine")],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        a = subprocess.run([os.path.join(tmp, "ref")], capture_output=True, text=True, timeout=60).stdout
        b = subprocess.run([os.path.join(tmp, "mine")], capture_output=True, text=True, timeout=60).stdout
        a = "\n".join(ln for ln in a.splitlines() if ln[:2] in ("P ", "C ", "B ", "G ", "S "))   # the reference also prints banners
        assert len(a.splitlines()) > 1000
        assert a == b.strip()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
