"""Install the UNMODIFIED reference package into baseline/_ref (git-ignored, travels to the GPU box with the snapshot) so
that bench.py's reference arm can also time the Python reference itself on the box's host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Needs /root/reference (build container).  The reference's build backend is
poetry-core, which is not in the image (`pip install /root/reference` fails with "No module named 'poetry'"), so the install
runs from a copy under /tmp whose pyproject.toml names setuptools as the backend - packaging metadata only, every file of
the `gym_multigrid` package is installed as it is in the checkout.  No dependencies are installed (gymnasium / matplotlib
are absent from the image and the wheelhouse): the package is imported under oracle/refshim, as the golden recorder does.

    python oracle/install_reference.py
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
TARGET = os.path.join(ROOT, "baseline", "_ref")

PYPROJECT = """[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"

[project]
name = "gym-multigrid"
version = "1.0.0"

[tool.setuptools.packages.find]
include = ["gym_multigrid*"]
"""


def install(force: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(REFERENCE, "gym_multigrid")):
        return None
    if os.path.isdir(os.path.join(TARGET, "gym_multigrid")) and not force:
        return TARGET
    tmp = "/tmp/mg_reference_copy"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(REFERENCE, tmp, ignore=shutil.ignore_patterns(".git", "__pycache__"))
    with open(os.path.join(tmp, "pyproject.toml"), "w") as f:
        f.write(PYPROJECT)
    shutil.rmtree(TARGET, ignore_errors=True)
    os.makedirs(TARGET, exist_ok=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
           "--target", TARGET, tmp]
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    shutil.rmtree(tmp, ignore_errors=True)
    if res.returncode != 0:
        raise RuntimeError("reference install failed:\n" + res.stdout + res.stderr)
    # the two map fixtures the reference's own tests use (tests/assets), for the CtF / Maze lines of the Python baseline
    assets = os.path.join(TARGET, "assets")
    os.makedirs(assets, exist_ok=True)
    for name in ("board.txt", "board_maze.txt"):
        src = os.path.join(REFERENCE, "tests", "assets", name)
        if os.path.exists(src):
            shutil.copy(src, os.path.join(assets, name))
    return TARGET


if __name__ == "__main__":
    print(install(force=True))
