"""Host-side pattern compiler of the product (find_tfbs_b200/csrc/tables.cpp) checked on CPU: the packed pair tables flag a window
iff its exact score exceeds min_score, in both field formats, for lengths 1..32, with N bases and several chunks (tests/cpp)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pair_tables_are_exact(tmp_path):
    exe = str(tmp_path / "test_tables")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_tables.cpp"),
                           os.path.join(ROOT, "find_tfbs_b200", "csrc", "tables.cpp")])
    p = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stdout
    assert "ALL OK" in p.stdout
