"""BASELINE config 5: k2trance.a2s `Song` started N times under the root voice
(44.1 kHz, 500-frame driver buffers as benchmark/benchmark.sh:50), rendered by
  a2render       the reference, CPU, one engine state (single-threaded by design)
  a2render_cuda  the UNMODIFIED reference host + our unit plug-in (drop-in mode)
on the same box; outputs compared bit for bit, a2_Run loop wall time reported.

    python profiles/cfg5_k2trance.py [copies] [frames] [--stats]

--stats sets A2CU_STATS=1: the plug-in times its own callbacks with rdtsc (adds ~15 % to the run).
`host_floor_s` is a third run with A2CU_NULL=1: the host's own VM + tree walk with no-op units.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import a2oracle as ao  # noqa: E402

STATS = "--stats" in sys.argv
argv = [a for a in sys.argv if a != "--stats"]
copies = int(argv[1]) if len(argv) > 1 else 1000
frames = int(argv[2]) if len(argv) > 2 else 88200
song = os.path.join(ao.REF_DIR, "songs", "benchmark", "k2trance.a2s")


def render(binary, null=False, dry=False):
    import subprocess
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".raw") as tf:
        env = dict(os.environ)
        if STATS and not null:
            env["A2CU_STATS"] = "1"
        if null:
            env["A2CU_NULL"] = "1"
        if dry:
            env["A2CU_DRY"] = "1"
        res = subprocess.run([os.path.join(ao.REF_DIR, binary), "-r", "44100", "-b", "500", "-n", str(frames),
                              "-x", str(copies), "-p", "Song", "-o", tf.name, os.path.basename(song)],
                             cwd=os.path.dirname(song), capture_output=True, text=True, env=env)
        if res.returncode:
            raise RuntimeError(res.stderr[-2000:])
        info = json.loads(res.stdout.strip().splitlines()[-1])
        info["stderr_tail"] = res.stderr.strip().splitlines()[-6:]
        return np.fromfile(tf.name, dtype="<i4").reshape(-1, 2), info


ref, ri = render("a2render")
out, oi = render("a2render_cuda")
# the unmodified host alone: same walk, same VM, every unit callback returns at once (plugin/a2cu_units.c
# A2CU_NULL) - the floor for ANY unit library behind the A2_unitdesc boundary
_, fi = render("a2render_cuda", null=True)
# recording only: the callbacks record as usual, the engine launches nothing and never waits (A2CU_DRY)
_, di = render("a2render_cuda", dry=True)
bad = np.nonzero((out != ref).any(axis=1))[0]
print(json.dumps({
    "workload": "cfg5: k2trance.a2s Song x %d, 44.1 kHz, buffer 500, %d frames" % (copies, frames),
    "bit_exact": bool(len(bad) == 0), "first_diff": int(bad[0]) if len(bad) else None,
    "active_voices_end": oi["active_voices"], "ref_active_voices_end": ri["active_voices"],
    "reference_cpu_s": ri["seconds"], "dropin_s": oi["seconds"],
    "host_floor_s": fi["seconds"], "host_plus_recording_s": di["seconds"], "speedup_ceiling_behind_the_boundary": ri["seconds"] / fi["seconds"],
    "speedup": ri["seconds"] / oi["seconds"], "realtime_factor_dropin": frames / 44100.0 / oi["seconds"],
    "peak": int(np.abs(ref).max()), "dropin_stats": oi["stderr_tail"]}))
