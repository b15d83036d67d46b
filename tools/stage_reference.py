#!/usr/bin/env python
"""BUILD CONTAINER ONLY: stage the unmodified reference under baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun)
with its Cython extensions rebuilt for this interpreter, so that tools/run_reference_main_gpu.py can drive the reference's own
main.py on a B200.  Nothing is edited; .ipynb_checkpoints debris is left out."""
import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = "/root/reference", os.path.join(REPO, "baseline", "_ref")

if __name__ == "__main__":
    if not os.path.isdir(SRC):
        sys.exit("no /root/reference here")
    shutil.rmtree(DST, ignore_errors=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns(".ipynb_checkpoints", "*.so"))
    subprocess.run(["chmod", "-R", "u+w", DST], check=True)
    subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=DST, check=True, stdout=subprocess.DEVNULL)
    print("staged", DST, [f for f in os.listdir(os.path.join(DST, "util", "cython")) if f.endswith(".so")])
