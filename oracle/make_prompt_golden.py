"""Golden vectors of the prompt simulators from the UNMODIFIED reference (build container only):
    python -m oracle.make_prompt_golden   ->  tests/golden/prompt_simulators.npz
Inputs are tests/test_prompts_cpu.py: make_case(seed); seeds random / np.random = seed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402


def main():
    rh.import_reference()
    from isegm.engine.trainer import get_next_promts
    from tests.test_prompts_cpu import SETTINGS, run
    out = {}
    for seed in range(6):
        for k, kw in enumerate(SETTINGS):
            for name, a in zip(("points", "boxes", "scribbles", "rects"), run(get_next_promts, seed, **kw)):
                out["s%d_k%d_%s" % (seed, k, name)] = a
    path = os.path.join(ROOT, "tests", "golden", "prompt_simulators.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
