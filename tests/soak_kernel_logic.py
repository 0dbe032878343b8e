#!/usr/bin/env python
"""Soak run (not collected by pytest): the host build of the kernel logic (tests/hostsim) against the C oracle over
millions of steps of randomised option sets — every step's observation, reward and done, hidden state every 50th.
A case stops when the kernel flags a full wolf table (counted by WAB_STAT_OVERFLOWS; results may differ from there).

    python tests/soak_kernel_logic.py SECONDS        # prints one JSON line per finished case
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.wab_oracle import OracleEnv
from tests.hostsim import HostSimEnv
from tests.test_fuzz_options import random_options
from tests.util import pick_action, mask_words_to_int, window_mask_from_bushes
total = 0; t0 = time.time(); case = 5000
while time.time() - t0 < float(sys.argv[1]):
    rng = np.random.default_rng(case)
    opts = random_options(rng) if case % 2 else {}
    auto = bool(case % 3)
    orc = OracleEnv(opts, seed=case, env_id=case * 31)
    sim = HostSimEnv(opts, seed=case, env_id=case * 31, auto_reset=auto, force_f64_food=not auto)
    oo = orc.reset(); so, _ = sim.reset(); done = False
    for n in range(20000):
        if done and not auto:
            oo = orc.reset(); so, _ = sim.reset(); done = False
        assert np.array_equal(oo[0], so[0]) and oo[1:] == so[1:], (case, n, opts)
        a = pick_action(rng, oo[0], orc.n_actions, greedy=bool(case & 1))
        oo, orr, od = orc.step(a)
        so, sr, sd, info, ovf = sim.step(a)
        if ovf: break          # wolf table full: counted by the kernel, results may differ from here on
        assert np.float32(orr) == np.float32(sr) and od == sd, (case, n, opts)
        if auto:
            if od: oo = orc.reset()
        else:
            done = od
        if n % 50 == 0 and not ovf:
            ho, hs = orc.hidden_state(), sim.hidden_state()
            assert (ho["x"], ho["y"], ho["turn"], ho["wolves"]) == (hs["x"], hs["y"], hs["turn"], hs["wolves"]), (case, n)
            assert mask_words_to_int(hs["bush_mask"]) == window_mask_from_bushes(ho), (case, n)
    total += n + 1; case += 1
    print(json.dumps({"cases": case - 5000, "steps": total, "seconds": round(time.time() - t0)}), flush=True)
