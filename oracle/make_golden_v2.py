"""Generate ``tests/golden/v2_trace.npz`` from the REAL Environment-2.0 reference (TEST INFRASTRUCTURE).

    python -m oracle.make_golden_v2

Runs "/root/reference/Environment 2.0" unmodified under ``oracle/ref_shim/v2.py`` (stub gym, keyed
``random.randint``) for a few worlds and records, per entity action, the observation list (as one-hot
planes + row count), internal observation, reward, done, and the full entity table after every turn."""
import os
import warnings

import numpy as np


def main():
    from tests.test_v2 import GOLDEN, WORLDS, record_reference
    episodes, turns = 3, 10
    out = {"n_worlds": np.int64(len(WORLDS)), "worlds": np.asarray(WORLDS, dtype=np.int64), "episodes": np.int64(episodes),
           "turns": np.int64(turns)}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for w, world in enumerate(WORLDS):
            rec = record_reference(world, episodes, turns)
            for k, v in rec.items():
                out["w%d_%s" % (w, k)] = v
            print("v2 golden world", world, "events", len(rec["reward"]), "kills", int((rec["state"][-1][:, 8] == 2).sum()), flush=True)
    os.makedirs(os.path.dirname(GOLDEN), exist_ok=True)
    np.savez_compressed(GOLDEN, **out)


if __name__ == "__main__":
    main()
