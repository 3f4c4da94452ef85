"""CPU: include/afmg.h is consumed by a plain C99 program (what a cgo / ISO_C_BINDING / JNI binding sees):
it must compile with -pedantic -Werror, link against libafmg.so and run its host-only checks."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c99_consumer_compiles_links_and_runs(tmp_path):
    lib_dir = os.path.join(ROOT, "afivo_streamer_b200")
    assert os.path.exists(os.path.join(lib_dir, "libafmg.so")), "build first: python __graft_entry__.py"
    exe = str(tmp_path / "c_abi_driver")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_driver.c"), "-o", exe, "-L", lib_dir, "-lafmg",
                           "-Wl,-rpath," + lib_dir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c abi ok" in out.stdout
