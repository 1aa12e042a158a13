"""Stage the UNMODIFIED reference package for the GPU box: /root/reference/mineral -> oracle/_ref/mineral.

TEST / BENCH INFRASTRUCTURE ONLY.  The reference (jacr13/mineral) is pure Python, so "building" it is what
`pip install --target` would do: place its package directory on a path.  The copy lives ONLY under `oracle/_ref/`,
which is listed in .gitignore (never in history) but not in .gpurunignore, so it travels to the GPU box with the
snapshot exactly like the built `.so` files do; `/root/reference` itself does not exist there.

Users: `bench.py --impl reference` (the reference's own `PPO` agent timed on the box's host cores, `cpu_baseline.kind`
"reference") and the optional live cross-checks of tests/test_oracle_golden.py.  Nothing under `mineral_b200/` reads it.

    python oracle/build_ref.py        # run by __graft_entry__.build() whenever /root/reference is present
"""

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MINERAL_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build(verbose=True):
    src_pkg = os.path.join(SRC, "mineral")
    if not os.path.isdir(src_pkg):
        if verbose:
            print(f"build_ref: {src_pkg} not present (GPU box): keeping whatever oracle/_ref holds")
        return os.path.isdir(os.path.join(DST, "mineral"))
    dst_pkg = os.path.join(DST, "mineral")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src_pkg, dst_pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST, "SOURCE.txt"), "w") as f:
        f.write(f"unmodified copy of {src_pkg} made by oracle/build_ref.py (git-ignored; not product source)\n")
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(dst_pkg))
        print(f"build_ref: staged {n} files under {dst_pkg}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
