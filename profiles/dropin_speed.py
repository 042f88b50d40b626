"""Drop-in throughput: the reference host + our plug-in (a2render_cuda) vs the
full reference (a2render) on the bench bank, same script, same box."""
import json, os, subprocess, sys, tempfile
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from cases import bench_bank
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 48000
scn = bench_bank(nv, steps=(frames + 959) // 960)
path = os.path.join(tempfile.mkdtemp(), "bank.a2s")
open(path, "w").write(scn.to_a2s())
for exe in ("a2render", "a2render_cuda"):
    out = subprocess.run([os.path.join("oracle/_ref", exe), "-n", str(frames), "-b", "64", "-p", "Song", path],
                         capture_output=True, text=True)
    info = json.loads(out.stdout.strip().splitlines()[-1])
    print("%-14s %d voices x %d frames: %.3f s in a2_Run -> %.1f M voice-samples/s (rt_error %d)" % (
        exe, nv, frames, info["seconds"], nv * frames / info["seconds"] / 1e6, info["rt_error"]))
