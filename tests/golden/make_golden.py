"""Mints the golden vectors under tests/golden/ from the UNMODIFIED reference (oracle/_ref).

Run here (where /root/reference exists):   python tests/golden/make_golden.py
For each scene it advances the reference to a step whose pressure solve iterates, then records the state after
every phase of that time_step (see pinlib.record_step).  The vectors travel to the GPU box; the reference
sources do not.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pinlib import GOLDEN_DIR, SCENES, make_scene, record_step  # noqa: E402

N = 10
WARM = {"cube_drop": 26, "dam_break": 3, "flip_obstacle": 4, "pic_sphere": 6}
DT = 0.004

if __name__ == "__main__":
    for scene in SCENES:
        ref = make_scene(scene, N)
        for _ in range(WARM[scene]):
            ref.time_step(DT)
        rec = record_step(ref, DT)
        meta = dict(size=np.array(ref.size, dtype=np.uint64), h=np.float64(ref.h), offset=ref.offset,
                    gravity=ref.gravity, method=np.int32(ref.method), blend=np.float64(ref.blend))
        path = os.path.join(GOLDEN_DIR, scene + ".npz")
        np.savez_compressed(path, **{"meta/" + k: v for k, v in meta.items()}, **rec)
        print(scene, ref.num_particles(), "iters", int(rec["solve/iters"]), os.path.getsize(path) // 1024, "KiB")
