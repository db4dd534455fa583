"""Oracle vs the reference's real stepper (oracle/_ref/libedyn_stepper.so) in free-running lock step at benchmark sizes.
Development aid / evidence generator: the CPU test-suite runs the same comparison on smaller scenes
(tests/test_ref_stepper.py); this takes a few minutes.

    python tools/ref_lockstep.py            # config 2 and config 3 at full size, slices of configs 4 and 5
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import edyn_b200 as E  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.test_ref_stepper import lockstep  # noqa: E402

RUNS = [("boxes_4096 (config 2, full size)", lambda: E.scenes.boxes_on_plane(16), 200),
        ("mixed_4096 (1/64 of config 4)", lambda: E.scenes.mixed_pile(16), 120),
        ("spheres_65536 (config 3, full size)", lambda: E.scenes.spheres_in_box(), 40),
        ("chains_65536 (1/16 of config 5)", lambda: E.scenes.hinge_chains(128, 128), 25)]

if __name__ == "__main__":
    O.build()
    assert O.ref_stepper() is not None, "make -C oracle stepper"
    for name, make, steps in RUNS:
        t0 = time.perf_counter()
        scene = make()
        first_bad, coarser, points = lockstep(O, scene, steps, threads=os.cpu_count() or 1)
        print(f"{name}: {scene['dynamic']} bodies, {steps} free-running steps, {points} contact points at the end: "
              f"{'bit-identical every step' if first_bad is None else f'first difference at step {first_bad}'}; "
              f"reference partition coarser on {coarser} steps; {time.perf_counter() - t0:.0f} s", flush=True)
