"""Compile the reference's OWN GoogleTest sources, unmodified and in place under /root/reference, against this
repo's C++ facade (include/) + libspx.so through tests/cpp/gtest_shim.  Outputs go to tests/cpp/_build/ref_<name>
(git-ignored; it travels to the GPU box, where /root/reference does not exist).  Test infrastructure only: no
reference source is copied into the repo, and nothing here is on a product path."""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_TESTS = "/root/reference/cpp/tests"
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
LIBDIR = os.path.join(ROOT, "sycl_points_b200")
# the reference test files whose whole API surface is inside this repo's scope
SOURCES = ["test_kdtree", "test_registration_pipeline", "test_downsampling_filters", "test_voxel_hash_map", "test_relative_pose_deskew", "test_preprocess_filter", "test_octree"]


# the reference's own example programs whose API surface is inside the scope: compiled the same way (no gtest)
REF_EXAMPLES = "/root/reference/cpp/examples"
EXAMPLES = ["example_registration", "example_point_cloud"]


def exe_path(name):
    return os.path.join(BUILD, "ref_" + name)


def build(verbose=False):
    """-> list of built executables ([] when the reference tree is not present)"""
    if not os.path.isdir(REF_TESTS):
        return []
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    os.makedirs(BUILD, exist_ok=True)
    shim = os.path.join(ROOT, "tests", "cpp", "gtest_shim")
    deps = [os.path.join(LIBDIR, "libspx.so"), os.path.join(shim, "gtest", "gtest.h")] + [
        os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(ROOT, "include")) for f in fs]
    newest = max(os.path.getmtime(p) for p in deps)
    out = []
    for name in SOURCES:
        src = os.path.join(REF_TESTS, name + ".cpp")
        exe = exe_path(name)
        if not (os.path.exists(exe) and os.path.getmtime(exe) > max(newest, os.path.getmtime(src))):
            cmd = [gxx, "-std=c++20", "-O1", "-I" + shim, "-I" + os.path.join(ROOT, "include"), src,
                   os.path.join(shim, "gtest_main.cpp"), "-L" + LIBDIR, "-lspx", "-Wl,-rpath," + LIBDIR, "-o", exe]
            env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            if r.returncode != 0:
                raise RuntimeError(f"{name}: {r.stderr[-3000:]}")
            if verbose:
                print("built", exe)
        out.append(exe)
    for name in EXAMPLES:
        src = os.path.join(REF_EXAMPLES, name + ".cpp")
        exe = exe_path(name)
        if not (os.path.exists(exe) and os.path.getmtime(exe) > max(newest, os.path.getmtime(src))):
            cmd = [gxx, "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"), src, "-L" + LIBDIR, "-lspx",
                   "-Wl,-rpath," + LIBDIR, "-o", exe]
            env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            if r.returncode != 0:
                raise RuntimeError(f"{name}: {r.stderr[-3000:]}")
            if verbose:
                print("built", exe)
        out.append(exe)
    return out


if __name__ == "__main__":
    build(verbose=True)
