"""Diagnostic table for tests/test_songs.py (run on the GPU box):
python tests/song_report.py [name-substring ...]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import test_songs as ts  # noqa: E402

rows = []
for where, name, program, frames, buffer in ts.BENCH + ts.TESTDATA:
    if len(sys.argv) > 1 and not any(a in name for a in sys.argv[1:]):
        continue
    path = os.path.join(ts.SONGS, where, name)
    ref, ri = ts._render(path, program, frames, buffer, "a2render")
    t0 = time.time()
    try:
        out, oi = ts._render(path, program, frames, buffer, "a2render_cuda")
    except Exception as e:  # noqa: BLE001
        rows.append({"song": where + "/" + name, "error": str(e)[-300:]})
        print(rows[-1], flush=True)
        continue
    bad = np.nonzero((out != ref).any(axis=1))[0]
    rows.append({"song": where + "/" + name, "ok": bool(len(bad) == 0 and oi["rt_error"] == ri["rt_error"]),
                 "rt_error": oi["rt_error"], "ref_rt_error": ri["rt_error"],
                 "voices": oi["active_voices"], "ref_voices": ri["active_voices"],
                 "first_diff": int(bad[0]) if len(bad) else None, "ndiff": int(len(bad)),
                 "maxabs": int(np.abs(out.astype(np.int64) - ref).max()),
                 "ref_s": ri["seconds"], "cuda_s": oi["seconds"], "wall": time.time() - t0})
    print(json.dumps(rows[-1]), flush=True)
