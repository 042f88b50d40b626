"""cfg2 (wtosc -> filter12 -> panmix) throughput against the number of voices in ONE launch:
what the engine picks by itself, render_split with one or two voice sets per CTA as the engine would
choose (A2CU_SPLIT_ALWAYS=1), render_split forced to one set (A2CU_ONE_SET=1), and the thread-per-voice
render_bank. 960-frame windows, kernel spans.

    python profiles/saturation.py > profiles/r02_saturation.json
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, '.')

if len(sys.argv) > 1 and sys.argv[1] == "--one":
    import numpy as np
    from audiality2_b200 import engine as eng
    from audiality2_b200.workloads import setup_cfg2
    V, split = int(sys.argv[2]), sys.argv[3] == "split"
    e = eng.Engine(48000, 2)
    e.set_split(split)
    bank, b = setup_cfg2(e, V)
    e.set_timing(True)
    ms = []
    for i in range(10):
        e.write_all(bank, 0, 2, [b['amp'] // (1 + i % 2)], dur=960 << 8)
        e.run(960, 64)
        if i >= 3:
            ms.append(e.last_render_ms() + e.last_mix_ms())
    print(json.dumps({"voices": V, "ms": float(np.median(ms)), "G_voice_samples_per_s": V * 960 / np.median(ms) / 1e6,
                      "kernel": ("render_split" if e.split_launches else "render_bank")}))
    e.close()
    sys.exit(0)

rows = []
for V in (4096, 8192, 16384, 32768, 65536, 131072, 262144):
    row = {"voices": V}
    for name, mode, env in (("engine_choice", "split", {}), ("split_auto", "split", {"A2CU_SPLIT_ALWAYS": "1"}),
                            ("split_one_set", "split", {"A2CU_ONE_SET": "1", "A2CU_SPLIT_ALWAYS": "1"}),
                            ("render_bank", "bank", {})):
        out = subprocess.run([sys.executable, __file__, "--one", str(V), mode], capture_output=True, text=True,
                             env=dict(os.environ, **env))
        r = json.loads(out.stdout.strip().splitlines()[-1])
        row[name] = {"ms": round(r["ms"], 4), "G": round(r["G_voice_samples_per_s"], 1)}
    rows.append(row)
    print(json.dumps(row), flush=True)
