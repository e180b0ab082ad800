"""Build the CPU checkers (TEST INFRASTRUCTURE; nothing in the product knows about this directory):
the plain-C restatement always, the reference-source build only where /root/reference exists (the build
container; the resulting files under oracle/_ref travel to the GPU box). See oracle/Makefile."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False):
    out = None if verbose else subprocess.DEVNULL
    if os.path.exists(os.path.join(HERE, "trimem_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=out)
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", HERE, "-j8", "ref"], stdout=out)


if __name__ == "__main__":
    build(verbose=True)
