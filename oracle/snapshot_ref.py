"""TEST / BENCH INFRASTRUCTURE ONLY -- snapshot of the unmodified reference for the CPU arm of bench.py.

The reference (DRL-Navigation/DDRL4NAV) is pure Python with no setup.py / pyproject.toml, so there is nothing to compile
and nothing pip can install.  What `bench.py --impl reference` needs on the GPU box -- where /root/reference does not
exist -- is the reference's own `USTC_lab` package importable.  `__graft_entry__.build()` calls `snapshot()` in the build
container: the package's *.py files are copied byte for byte into oracle/_ref/ (git-ignored, NOT gpurun-ignored: it
travels with the repo snapshot exactly like a built .so; it never enters the history).  Nothing under ddrl4nav_b200/
imports it; only bench.py's CPU legs and the `not gpu` cross-checks do (through oracle/ref_shim.py)."""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def snapshot(src_root: str = "/root/reference") -> bool:
    src = os.path.join(src_root, "USTC_lab")
    if not os.path.isdir(src):
        return os.path.isdir(os.path.join(DST, "USTC_lab"))
    dst = os.path.join(DST, "USTC_lab")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    n = 0
    for d, _, files in os.walk(src):
        rel = os.path.relpath(d, src)
        for f in files:
            if f.endswith(".py"):
                os.makedirs(os.path.join(dst, rel), exist_ok=True)
                shutil.copyfile(os.path.join(d, f), os.path.join(dst, rel, f))
                n += 1
    with open(os.path.join(DST, "README"), "w") as fh:
        fh.write("snapshot of %s (*.py only, %d files) made by oracle/snapshot_ref.py; git-ignored\n" % (src, n))
    return True


if __name__ == "__main__":
    print("snapshot:", snapshot())
