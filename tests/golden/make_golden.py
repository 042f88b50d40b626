"""Generate tests/golden/ref_outputs.npz from the REFERENCE itself.

Run in the dev container (needs oracle/_ref, i.e. /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

For every scenario in tests/cases.py the generated .a2s script is rendered by
oracle/_ref/a2render (the unmodified reference, `buffer` driver) and the int32
8:24 master output is stored.  The reference has no golden vectors of its own
(SURVEY.md section 4), so these files are the pin for oracle/a2_oracle.c and,
through it, for the CUDA path.  The scripts are stored too (tests/golden/*.a2s)
so a reader can see exactly what the reference was asked to play.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from cases import CASES, SCRIPTS   # noqa: E402
from scenarios import run_ref      # noqa: E402
from oracle import a2oracle as ao  # noqa: E402


def main():
    out = {}
    for name, build in CASES.items():
        scn = build()
        path = os.path.join(HERE, name + ".a2s")
        data = run_ref(scn, path)
        out[name] = data
        print("%-24s %s frames=%d sha256=%s" % (
            name, data.shape, scn.frames,
            hashlib.sha256(data.tobytes()).hexdigest()[:16]))
    # hand-written scripts (tests/data/*.a2s): whole songs through the reference
    for name, (path, program, frames, rate, buffer) in SCRIPTS.items():
        data, info = ao.ref_render(os.path.join(os.path.dirname(HERE), path), program,
                                   samplerate=rate, channels=2, buffer=buffer, frames=frames)
        assert info["rt_error"] == 0
        out["script_" + name] = data
        print("%-24s %s frames=%d sha256=%s" % (
            "script_" + name, data.shape, frames, hashlib.sha256(data.tobytes()).hexdigest()[:16]))
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)


if __name__ == "__main__":
    main()
