"""Drop-in check at the source level: the reference's UNMODIFIED ROS 2 node sources (main_path_planning.cpp of
reeds_shepp_hybrid_astar and dubins_hybrid_star) compile against the hpp-b200 host shells when the package's five class
headers are replaced by the forwarding headers of hybrid_path_planning_b200/host/dropin/ (INTEGRATION.md section 1).  ROS itself is not installed here: tests/ros_mock
holds stand-ins for the rclcpp / message / tf2 headers the nodes include.  Needs the reference checkout (skipped on
machines without it, e.g. the GPU box); nothing of the reference is copied into the repository -- its node header goes to
a temporary directory for the duration of the compile."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HPP_REF", "/root/reference")
MOCK = os.path.join(ROOT, "tests", "ros_mock")


@pytest.mark.parametrize("tag,pkg", [("rs", "reeds_shepp_hybrid_astar"), ("db", "dubins_hybrid_star")])
def test_reference_node_compiles_on_the_shells(tag, pkg):
    src = os.path.join(REF, pkg, "src", "main_path_planning.cpp")
    if not os.path.exists(src):
        pytest.skip("reference checkout not present")
    tmp = tempfile.mkdtemp()
    try:
        fwd = os.path.join(ROOT, "hybrid_path_planning_b200", "host", "dropin", tag)
        for f in os.listdir(fwd):
            shutil.copy(os.path.join(fwd, f), tmp)
        shutil.copy(os.path.join(REF, pkg, "include", "main_path_planning.h"), tmp)
        cmd = ["g++", "-std=c++17", "-fsyntax-only", "-w", "-I" + tmp, "-I" + os.path.join(MOCK, "include"),
               "-I" + os.path.join(ROOT, "hybrid_path_planning_b200", "host"), "-I" + os.path.join(ROOT, "include"), src]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def test_shell_headers_stand_alone():
    """the shells compile without any ROS header on the include path (CarData's outline falls back to its own Polygon)"""
    code = '#include "HybridAstarShell.h"\nint main() { CarData c(0.6, 2.75, 1.1, 1.2, 1.9); c.createVehicleGeometry(); ' \
           'return (int)c.getVehicleGeometry_state(State(1, 2, 0.5)).points.size() + (int)c.getCircles(State(0, 0, 0)).size() - 9; }\n'
    tmp = tempfile.mkdtemp()
    try:
        with open(os.path.join(tmp, "t.cpp"), "w") as f:
            f.write(code)
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "hybrid_path_planning_b200", "host"),
                            "-I" + os.path.join(ROOT, "include"), os.path.join(tmp, "t.cpp")], capture_output=True, text=True,
                           timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def test_reference_map_inflate_node_compiles_on_the_cv_shim():
    """MI/src/map_inflate_node.cpp, unmodified, against hybrid_path_planning_b200/host/dropin/mi (opencv2/opencv.hpp and
    cv_bridge/cv_bridge.h stand-ins that route cv::threshold / cv::dilate to the C ABI) and the stand-in rclcpp"""
    src = os.path.join(REF, "map_inflate", "src", "map_inflate_node.cpp")
    if not os.path.exists(src):
        pytest.skip("reference checkout not present")
    host = os.path.join(ROOT, "hybrid_path_planning_b200", "host")
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-w", "-I" + os.path.join(host, "dropin", "mi"),
           "-I" + os.path.join(MOCK, "include"), "-I" + host, "-I" + os.path.join(ROOT, "include"), src]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.parametrize("tag,hdr,code", [
    ("rs", "reeds_shepp_path.h", "reeds_shepp_path p; PATH q = p.calc_optimal_path(0, 0, 0, 5, 3, 1, 0.25); "
                                 "std::vector<PATH> a = p.calc_all_paths(0, 0, 0, 5, 3, 1, 0.25, STEP_SIZE); "
                                 "return (int)a.size() + (int)q.directions.size() + (int)q.ctypes.size();"),
    ("db", "dubins_path.h", "dubins_path p; Path q = p.calc_dubins_path(0, 0, 0, 5, 3, 1, 0.25); "
                            "return (int)q.mode.size() + (int)q.x.size() + (int)p.mod2pi(7.0) + (int)p.pi_2_pi(4.0);")])
def test_path_class_forwarding_headers(tag, hdr, code):
    """code written against the reference's reeds_shepp_path / dubins_path headers compiles against the forwarding
    headers of host/dropin (same class names, value types, signatures and default arguments)"""
    tmp = tempfile.mkdtemp()
    try:
        with open(os.path.join(tmp, "t.cpp"), "w") as f:
            f.write(f'#include "{hdr}"\nint main() {{ {code} }}\n')
        host = os.path.join(ROOT, "hybrid_path_planning_b200", "host")
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(host, "dropin", tag), "-I" + host,
                            "-I" + os.path.join(ROOT, "include"), os.path.join(tmp, "t.cpp")], capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
