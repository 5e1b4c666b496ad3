"""In-tree build of libb2r.so: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo (see
csrc/Makefile).  Cross-compiles without a GPU in seconds."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False, jobs=None):
    jobs = jobs or os.cpu_count() or 4
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j%d" % jobs]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return os.path.join(_HERE, "libb2r.so")


if __name__ == "__main__":
    print(build(verbose=True))
