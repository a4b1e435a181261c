"""Build the CPU replay library used by the non-GPU tests (test infrastructure only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhopb_emul.so")


def build(force=False):
    src = os.path.join(HERE, "emul.cpp")
    hdr_dir = os.path.join(HERE, "..", "..", "hop_b200", "csrc")
    deps = [src] + [os.path.join(hdr_dir, h) for h in ("hopb_binning.h", "hopb_machine.h", "hopb_gridtab_host.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", SO, src])
    return SO


if __name__ == "__main__":
    print(build(force=True))
