#!/usr/bin/env python
"""End-to-end error of the benched math mode (and of the fp32 class) against the CPU oracle at full size for the workloads the
test suite does not cover by default (the suite runs dtu): calls tests/test_network_gpu.py::test_benched_math_mode_against_oracle,
which writes gpurun_out/benched_mode_parity_<workload>.json before it asserts.  python tools/benched_parity.py [nerf llff]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_network_gpu as T

for w in sys.argv[1:] or ["nerf", "llff"]:
    try:
        T.test_benched_math_mode_against_oracle(w)
        print(w, "within the test's bounds", flush=True)
    except AssertionError as exc:
        print(w, "outside the test's bounds:", str(exc)[:300], flush=True)
