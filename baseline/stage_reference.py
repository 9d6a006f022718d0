"""Stage the UNMODIFIED reference under baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box).

    python baseline/stage_reference.py [--force]

The reference is four Python files with no build system (`pip install /root/reference` has nothing to install:
no setup.py / pyproject.toml), so "installing" it is copying its files: optex.py, histmatch.py, util.py, vgg.py, the
trained weights models/*.pth and the bundled images style/*.jpg, content/*.jpg.  Nothing here is imported by the
product; bench.py --impl reference, the cpu_baseline leg and tests/ use it through baseline/reference.py.
Runs only where /root/reference exists (the build container); a no-op elsewhere.
"""
import os
import shutil
import sys

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["optex.py", "histmatch.py", "util.py", "vgg.py", "requirements.txt"]
DIRS = ["models", "style", "content"]


def stage(force: bool = False) -> str:
    if not os.path.isdir(SRC):
        return "skipped: /root/reference is not present on this machine"
    os.makedirs(DST, exist_ok=True)
    copied = 0
    for f in FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        if os.path.exists(s) and (force or not os.path.exists(d) or os.path.getsize(s) != os.path.getsize(d)):
            shutil.copyfile(s, d)
            copied += 1
    for sub in DIRS:
        os.makedirs(os.path.join(DST, sub), exist_ok=True)
        for f in sorted(os.listdir(os.path.join(SRC, sub))):
            s, d = os.path.join(SRC, sub, f), os.path.join(DST, sub, f)
            if os.path.isfile(s) and (force or not os.path.exists(d) or os.path.getsize(s) != os.path.getsize(d)):
                shutil.copyfile(s, d)
                copied += 1
    return f"staged {copied} file(s) into {DST}"


if __name__ == "__main__":
    print(stage("--force" in sys.argv))
