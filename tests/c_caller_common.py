"""Builds and runs tests/c/battle_caller.c (a plain-C caller of the reference ABI) -- test infrastructure."""
import os
import subprocess

import parity_common as pc

SRC = os.path.join(pc.REPO, "tests", "c", "battle_caller.c")
BIN = os.path.join(pc.REPO, "tests", "_c", "battle_caller")


def build():
    if os.path.exists(BIN) and os.path.getmtime(BIN) >= os.path.getmtime(SRC):
        return BIN
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    # -std=c99 -pedantic-errors: the headers must be plain C (no C++/CUDA/torch types cross the boundary)
    subprocess.run(["gcc", "-std=c99", "-pedantic-errors", "-D_DEFAULT_SOURCE", "-O1", "-Wall", "-Wextra", "-Werror",
                    "-I", os.path.join(pc.REPO, "include"), SRC, "-o", BIN, "-ldl"], check=True)
    return BIN


def run(lib, steps=40, size=40, n=250):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run([build(), lib, str(steps), str(size), str(n)], capture_output=True, text=True, env=env,
                          timeout=300)
