"""The measurement scripts under tools/ run only on the GPU box; a typo there costs a whole GPU call.  Compile every
Python tool, syntax-check the shell scripts, and import-check the ones whose imports do not need CUDA at import time."""
import glob
import os
import py_compile
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tools", "*.py")) + [os.path.join(ROOT, "bench.py"),
                                                                                          os.path.join(ROOT, "__graft_entry__.py")]))
def test_python_tool_compiles(path):
    py_compile.compile(path, doraise=True)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tools", "*.sh")) + glob.glob(os.path.join(ROOT, "tests", "*.sh"))))
def test_shell_script_parses(path):
    if shutil.which("bash") is None:
        pytest.skip("no bash")
    subprocess.run(["bash", "-n", path], check=True)
