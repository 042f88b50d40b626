#!/usr/bin/env python
"""bench.py - voice-samples/s of the voice-render hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path

Workload (config.workload "cfg2"): 4 096 voices per GPU, wtosc -> filter12 -> panmix on
the shared 2 048-point saw, 48 kHz, 64-frame blocks, plus ONE control write per voice
per 20 ms (amplitude re-targeted and ramped) so every window has real host->device input.
One WINDOW = 960 frames (20 ms = 15 blocks of 64; one a2cu_submit, one kernel launch).
One STEP = 50 windows = 1 s of audio = voices x 48 000 voice-samples.

Numbers on the JSON line:
  value   voice-samples/s over the CUDA-event spans of the kernels (the window's events are
          already in HBM when a span starts), summed over the timed windows, max over
          ranks. For N > 1 the span includes the in-kernel NVLink exchange (the wait for
          the peers' root buses). No L2 flush: every window renders a DIFFERENT bank of
          4096 voices, round-robin over enough banks that their state exceeds the L2 by
          1.5x (inputs larger than L2).
  e2e     the same metric through the public C ABI with HOST buffers: per window
          a2cu_bank_write_all (host events) -> a2cu_submit (event staging, H2D, kernel; the
          root stage writes the int32 master block into pinned host memory) ->
          a2cu_collect (wait + copy to the caller's buffer), two windows in flight; wall
          time of the whole timed region, max over ranks. Same code path for every N.
  roofline  HBM roofline of the dominant kernel (render_split<...>, ONE launch per window:
          the root stage - and for N > 1 the root-bus exchange - is fused into its last CTA).
  configs   further workloads of BASELINE.json measured in the same run (kernel spans,
          L2 flushed between windows where the inputs fit the L2): cfg3 (65 536 additive
          voices), the cfg4 per-GPU shard (32 768 FM voices; for N > 1 sharded over the
          ranks with the exchange), the HBM-bound sampled-wave gather with ITS roofline,
          and cfg2 at 64- and 256-frame windows (the latency-honest lines).
  cpu_baseline  the reference's own CPU render (oracle/_ref) on a bounded sample of the
          same workload, 1 core, rank 0, N = 1 only.
Multi-GPU (weak scaling): every rank renders its own voices; the raw stereo root bus is
summed INSIDE the render kernel over NVLink peer memory (a2cu_xchg_*, csrc/a2cu_kernels.cuh
xchg_root_bus) before the truncating root stage. `--exchange nccl` runs the library
baseline instead (cut path + NCCL all-reduce + separate root-stage launch) for A/B.

The GPU arm imports nothing from tests/ or oracle/; the reference arm and the
cpu_baseline leg (test infrastructure by definition) generate the workload's .a2s script
with tests/cases.bench_bank and run oracle/_ref.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STEP_MS = 20
RATE = 48000
WINDOW = STEP_MS * RATE // 1000         # 960 frames = 15 blocks of 64
BLOCK = 64
WINDOWS_PER_STEP = 50                   # one step = 1 s of audio
L2_BYTES = 126 * 1024 * 1024            # B200 L2
NCU_SUMMARY = os.path.join("profiles", "r02b_render_split_ncu.txt")
NCU_FALLBACK = os.path.join("profiles", "r02_render_split_ncu.txt")


def workload_name(voices):
    return ("cfg2: %d voices per GPU wtosc->filter12->panmix, saw 2048-pt, 48 kHz, 64-frame blocks, "
            "1 amplitude write/voice/20 ms" % voices)


def config_of(voices):
    """Identical in both arms (the driver compares them)."""
    return {
        "workload": workload_name(voices),
        "step": "%d windows of %d frames (15 blocks of 64) = 1 s of audio per step" % (WINDOWS_PER_STEP, WINDOW),
        "l2": "GPU arm: no flush, inputs larger than L2 - every window renders a different bank, round-robin "
              "over banks whose per-voice state totals 1.5x the 126 MB L2",
    }


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the
    committed `ncu --set full` summary of this workload (static: a profiler cannot run inside the
    timed region)."""
    for rel in (NCU_SUMMARY, NCU_FALLBACK):
        path = os.path.join(ROOT, rel)
        try:
            tot = 0.0
            for ln in open(path):
                f = ln.split()
                if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(f[2], 1)
                    tot += float(f[1]) * mult
            if tot:
                return tot, rel
        except Exception:
            pass
    return None, None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown",
                                  "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# reference / cpu baseline (oracle/_ref: the unmodified reference, CPU)
# ---------------------------------------------------------------------------
def _write_script(nvoices, steps, seed, path):
    tests = os.path.join(ROOT, "tests")
    if tests not in sys.path:
        sys.path.insert(0, tests)
    from cases import bench_bank
    scn = bench_bank(nvoices, steps=steps, step_ms=STEP_MS, seed=seed)
    with open(path, "w") as f:
        f.write(scn.to_a2s())


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "a2render")
    return p if os.path.exists(p) else None


def run_reference_sample(nvoices, frames, shards, seed=324357):
    """Render `frames` frames of the cfg2 bench bank with the reference on `shards` host
    processes (one engine state is single-threaded, so the only legal parallelism is independent
    states on disjoint voice shards, audiality2.h.cmake:163-166).
    Returns (voice_samples, seconds, kind, processes)."""
    exe = ref_binary()
    steps = (frames + WINDOW - 1) // WINDOW
    if exe is None:
        # plain-C port (single thread)
        tests = os.path.join(ROOT, "tests")
        if tests not in sys.path:
            sys.path.insert(0, tests)
        from cases import bench_bank
        from scenarios import run_oracle
        scn = bench_bank(nvoices, steps=steps, step_ms=STEP_MS, seed=seed)
        scn.frames = frames
        t0 = time.perf_counter()
        run_oracle(scn)
        return nvoices * frames, time.perf_counter() - t0, "port", 1
    per = (nvoices + shards - 1) // shards
    tmp = tempfile.mkdtemp(prefix="a2ref_")
    procs = []
    for s in range(shards):
        n = min(per, nvoices - s * per)
        if n <= 0:
            break
        path = os.path.join(tmp, "shard%d.a2s" % s)
        _write_script(n, steps, seed + s, path)
        procs.append(subprocess.Popen(
            [exe, "-r", str(RATE), "-b", str(BLOCK), "-c", "2", "-n", str(frames),
             "-p", "Song", path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    secs = []
    for p in procs:
        out, err = p.communicate()
        if p.returncode:
            raise RuntimeError("a2render failed: " + err[-500:])
        secs.append(json.loads(out.strip().splitlines()[-1])["seconds"])
    return nvoices * frames, max(secs), "reference", len(procs)


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames = WINDOW * WINDOWS_PER_STEP      # one step = 1 s of audio, as in the GPU arm
    times, used, kind = [], 1, "reference"
    for i in range(args.warmup + args.steps):
        vs, sec, kind, used = run_reference_sample(args.voices, frames, cores)
        if i >= args.warmup:
            times.append(sec)
    total = args.voices * frames * len(times)
    T = sum(times)
    value = total / T
    sample = "%d voices x %d frames per step, %d independent engine states on %d host threads" % (
        args.voices, frames, used, used)
    line = {
        "impl": "reference", "metric": "voice-samples/sec at 64-frame blocks",
        "value": value, "unit": "voice-samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * T / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config_of(args.voices),
        "cpu_baseline": {"value": value, "unit": "voice-samples/s", "cores": used,
                         "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "voice-samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
class Cfg2Runner:
    """The cfg2 pipeline through the public C ABI: round-robin banks, one amplitude write per
    voice and window, a2cu_submit / a2cu_collect with two windows in flight."""

    def __init__(self, e, banks, amp):
        import numpy as np
        self.e, self.L, self.banks = e, e.L, banks
        self.amp = [np.array([amp // 2], dtype=np.int32), np.array([amp], dtype=np.int32)]
        self.i = 0
        self.pending = []
        self.out = None
        self.np = np
        self.dev_ms, self.render_ms = [], []

    def window(self, frames, timed):
        e, L, n = self.e, self.L, len(self.banks)
        i = self.i
        cur = self.banks[i % n]
        L.a2cu_bank_enable(e.h, self.banks[(i - 1) % n], 0)      # pause the previous window's bank
        L.a2cu_bank_enable(e.h, cur, 1)
        a = self.amp[(i // n) & 1]
        L.a2cu_bank_write_all(e.h, cur, 0, 2, a.ctypes.data, 0, L.a2cu_now(e.h), frames << 8)
        self.pending.append((e.submit(frames, BLOCK), timed))
        if len(self.pending) > 2:
            self.collect_one()
        self.i += 1

    def collect_one(self):
        t, timed = self.pending.pop(0)
        frames = self.e._frames_of[t]
        if self.out is None or self.out.shape[0] != frames:
            self.out = self.np.empty((frames, 2), dtype=self.np.int32)
        self.e.collect(t, self.out)                 # waits for THAT window only
        if timed:
            self.dev_ms.append(self.e.last_render_ms() + self.e.last_mix_ms())
            self.render_ms.append(self.e.last_render_ms())

    def drain(self):
        while self.pending:
            self.collect_one()


def span_windows(e, frames, buffer, nwin, warm, flush=None):
    """Kernel-span time (ms, median) of one a2cu_run window of whatever banks `e` holds."""
    ms = []
    for k in range(warm + nwin):
        if flush is not None:
            flush.zero_()               # > L2, on the engine's stream: evicts the previous window's lines
        e.run(frames, buffer)
        if k >= warm:
            ms.append(e.last_render_ms() + e.last_mix_ms())
    return statistics.median(ms)


def secondary_configs(local, dev, world, rank, peak, args):
    """Other BASELINE.json workloads, measured after the main timed region (kernel spans)."""
    import numpy as np
    import torch
    from audiality2_b200 import engine as eng
    from audiality2_b200 import workloads as wl
    from audiality2_b200.parallel import connect_engines
    stream = torch.cuda.current_stream()
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = {}

    def fresh():
        e = eng.Engine(RATE, 2, device=local)
        e.set_stream(stream.cuda_stream)
        e.set_timing(True)
        return e

    if world == 1:
        # cfg3: 65 536 additive voices, 256-frame buffers
        e = fresh()
        banks, _ = wl.setup_cfg3(e, 65536)
        frames = 1024
        ms = span_windows(e, frames, 256, 8, 3, flush)
        sb = e.bank_state_bytes(banks[0])
        alg = 65536 * 2 * sb
        out["cfg3"] = {
            "workload": "65 536 voices 8 x wtosc + panmix, sine, 256-frame blocks, %d-frame windows" % frames,
            "value": 65536 * frames / (ms / 1e3), "unit": "voice-samples/s", "ms_per_window": ms,
            "kernel": "render_split<%s>" % e.bank_kernel_name(banks[0]) if e.split_launches else
                      "render_bank<%s>" % e.bank_kernel_name(banks[0]),
            "l2": "flushed between windows (192 MB write)",
            "roofline": {"bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg,
                         "note": "state read + written once per launch; bound by INT32 issue / smem gather"}}
        e.close()
        # the HBM-bound gather: large sampled waves, at 64 and at 32 wave samples per output frame
        for key, spf in (("gather", 64), ("gather32", 32)):
            e = fresh()
            V = args.gather_voices
            banks, info = wl.setup_gather(e, V, samples_per_frame=spf)
            frames = 256
            e.run(frames, 64)               # uploads the 403 MB wave pool
            ms = span_windows(e, frames, 64, 6, 1, None)
            alg = float(V) * frames * info["bytes_per_voice_sample"]
            note = ("2 Hermite taps per output sample, each in its own 32-byte sector: 64 B per voice-sample "
                    "(SURVEY.md 8(d)). Stage A runs lane = frame, so one warp instruction reads a contiguous run "
                    "of the wave; ")
            if spf == 64:
                note += ("at 64 samples per frame the taps are 64 B apart and only every other sector of the run "
                         "is used, while DRAM delivers 64-byte atoms: ncu counts 3.2 GB read for 2.1 GB "
                         "algorithmic (profiles/r02b_gather64_ncu.txt), i.e. DRAM itself runs at 1.5x this fraction")
            else:
                note += ("at 32 samples per frame every sector of the run is used: DRAM traffic = algorithmic "
                         "bytes less L2 hits (profiles/r02b_gather32_ncu.txt)")
            out[key] = {
                "workload": "%d voices wtosc->panmix on 12 sampled waves x 16.8 M samples (%.0f MB, 3x L2), "
                            "%d wave samples per frame, %d-frame windows" % (V, info["wave_bytes"] / 1e6, spf, frames),
                "value": V * frames / (ms / 1e3), "unit": "voice-samples/s", "ms_per_window": ms,
                "kernel": "%s<%s>" % ("render_split" if e.split_launches else "render_bank",
                                      e.bank_kernel_name(banks[0])),
                "l2": "no flush: inputs (wave pool) are 3x the L2",
                "roofline": {"bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": alg / (ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg,
                             "note": note}}
            e.close()
    if world == 1:
        # cfg2 at saturation: one bank of 262 144 voices in one launch (profiles/r02_saturation.json)
        e = fresh()
        V = 262144
        bank, b = wl.setup_cfg2(e, V, ramp_frames=WINDOW)
        ms_l = []
        for k in range(9):
            flush.zero_()
            e.write_all(bank, 0, 2, [b["amp"] // (1 + k % 2)], dur=WINDOW << 8)
            e.run(WINDOW, BLOCK)
            if k >= 3:
                ms_l.append(e.last_render_ms() + e.last_mix_ms())
        ms = statistics.median(ms_l)
        sb = e.bank_state_bytes(bank)
        alg = V * (2 * sb + 20)
        out["cfg2_262144"] = {
            "workload": "cfg2 with 262 144 voices in one bank (one launch per 960-frame window)",
            "value": V * WINDOW / (ms / 1e3), "unit": "voice-samples/s", "ms_per_window": ms,
            "kernel": "%s<%s>" % ("render_split" if e.split_launches else "render_bank", e.bank_kernel_name(bank)),
            "l2": "flushed between windows (192 MB write)",
            "roofline": {"bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg,
                         "note": "one thread per voice above ~43 k voices: bound by INT32 issue"}}
        e.close()
    # cfg4: 32 768 FM voices per GPU (BASELINE's named multi-GPU config: 262 144 voices over 8 GPUs)
    e = fresh()
    V = 32768
    banks = wl.setup_cfg4(e, V, first_voice=rank * V, total=world * V)
    frames = 960
    if world > 1:
        connect_engines(e, frames, timeout_ms=20000, device=dev)
    ms = span_windows(e, frames, 64, 8, 3, flush)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    sb = sum(e.bank_state_bytes(b) for b in banks) / len(banks)
    alg = V * 2 * sb
    out["cfg4_shard"] = {
        "workload": "%d voices (32 768 per GPU) cycling fm3 / fm3p / fm2r / fm4r + panmix, 64-frame blocks, "
                    "%d-frame windows%s" % (V * world, frames,
                                            ", root bus summed over NVLink inside the root-stage kernel"
                                            if world > 1 else ""),
        "value": float(V) * world * frames / (ms / 1e3), "unit": "voice-samples/s", "ms_per_window": ms,
        "kernel": "render_bank<fmX_panmix> x 4 + %s" % ("mix_root_xchg" if world > 1 else "mix_root"),
        "l2": "flushed between windows (192 MB write)",
        "roofline": {"bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg,
                     "note": "bound by INT32 issue (4x oversampled operators), not HBM"}}
    e.close()
    del flush
    return out


def bench_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from audiality2_b200 import engine as eng
    from audiality2_b200 import workloads as wl
    from audiality2_b200.parallel import connect_engines, reduce_root_bus

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    multi = world > 1
    fused = multi and args.exchange == "fused"

    e = eng.Engine(RATE, 2, device=local)
    stream = torch.cuda.current_stream()
    e.set_stream(stream.cuda_stream)
    e.set_timing(True)
    # Round-robin over R independent banks of `voices` voices (one per window), so that the
    # per-voice state a window reads was last touched R windows ago and the state of all banks
    # together (R x voices x state bytes) exceeds the 126 MB L2 by 1.5x.
    banks, params = [], None
    nbanks = args.banks
    while True:
        r = len(banks)
        seed = 324357 + rank + (1000 * r if r < 8 else 0)
        bank, b = wl.setup_cfg2(e, args.voices, seed=seed, ramp_frames=WINDOW)
        e.bank_enable(bank, False)
        banks.append(bank)
        params = params or b
        if nbanks <= 0:
            nbanks = int(1.5 * L2_BYTES / (e.bank_state_bytes(bank) * args.voices)) + 1
        if len(banks) >= nbanks:
            break
    state_mb = nbanks * e.bank_state_bytes(banks[0]) * args.voices / 1e6
    if fused:
        connect_engines(e, WINDOW, timeout_ms=20000, device=dev)
    run = Cfg2Runner(e, banks, params["amp"])

    # ---- parity of the sharded path, checked in THIS run: window 0 (bank 0 of every rank) against
    # ---- ONE engine on rank 0 rendering all ranks' voices
    parity = None
    if multi:
        if fused:
            run.window(WINDOW, False)
            run.drain()
            first = run.out.copy()
        else:
            first = nccl_window(e, run, dev, stream, reduce_root_bus, False).copy()
        if rank == 0:
            chk = eng.Engine(RATE, 2, device=local)
            for rr in range(world):
                cb, cp = wl.setup_cfg2(chk, args.voices, seed=324357 + rr, ramp_frames=WINDOW)
                chk.write_all(cb, 0, 2, [cp["amp"] // 2], dur=WINDOW << 8)
            whole = chk.run(WINDOW, BLOCK)
            chk.close()
            if not np.array_equal(first, whole):
                raise SystemExit("bench: sharded window 0 differs from the single-engine render of all shards")
            parity = "window 0 of %d ranks == one engine rendering all %d voices (bit-exact, peak %d)" % (
                world, world * args.voices, int(np.abs(whole).max()))

    def one_window(timed):
        if multi and not fused:
            nccl_window(e, run, dev, stream, reduce_root_bus, timed)
        else:
            run.window(WINDOW, timed)

    def drain():
        run.drain()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # >= max(3, W, nbanks) warm-up windows (every bank rendered once before timing), then chunks of
    # 50 more until the GPU has been under this load for ~1.5 s, so nvidia-smi (100 ms period)
    # samples clocks under load. Rank 0 decides and broadcasts (all ranks must render the same
    # number of windows: the exchange is collective).
    nwarm = max(3, args.warmup, nbanks)
    t_w = time.perf_counter()
    for _ in range(nwarm):
        one_window(False)
    drain()
    while True:
        go = torch.tensor([1 if time.perf_counter() - t_w < 1.5 and nwarm < 100000 else 0], device=dev)
        if multi:
            dist.broadcast(go, 0)
        if not int(go.item()):
            break
        for _ in range(50):
            one_window(False)
        drain()
        nwarm += 50
    if multi:
        dist.barrier()
    torch.cuda.synchronize()
    l0, h0, d0 = e.launches, e.h2d_bytes, e.d2h_bytes
    nwin = args.steps * WINDOWS_PER_STEP
    wall0 = time.perf_counter()
    for _ in range(nwin):
        one_window(True)
    drain()
    torch.cuda.synchronize()
    if multi:
        dist.barrier()
    wall1 = time.perf_counter()
    launches = e.launches - l0
    h2d = (e.h2d_bytes - h0) / args.steps
    d2h = (e.d2h_bytes - d0) / args.steps
    if multi and not fused:
        d2h = WINDOW * 2 * 4 * WINDOWS_PER_STEP
    clk = clocks.stop() if rank == 0 else None
    dev_ms = run.dev_ms if not (multi and not fused) else NCCL_STATE["dev_ms"]
    render_ms = run.render_ms if not (multi and not fused) else NCCL_STATE["render_ms"]

    tot = torch.tensor([sum(dev_ms), (wall1 - wall0) * 1000.0], dtype=torch.float64, device=dev)
    if multi:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    dev_total_ms, host_total_ms = [float(x) for x in tot.tolist()]
    vs_total = float(args.voices) * WINDOW * nwin * world
    value = vs_total / (dev_total_ms / 1000.0)
    e2e = vs_total / (host_total_ms / 1000.0)
    k_ms = statistics.mean(render_ms)
    kname = "render_split" if e.split_launches else "render_bank"
    kernel_name = e.bank_kernel_name(banks[0])
    state_bytes = e.bank_state_bytes(banks[0])

    # ---- cfg2 at short windows (latency-honest lines): same engine, same banks ----
    short = {}
    if not multi and not args.no_configs:
        for frames in (256, 64):
            r2 = Cfg2Runner(e, banks, params["amp"])
            r2.i = run.i
            for _ in range(60):
                r2.window(frames, False)
            r2.drain()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n2 = 600
            for _ in range(n2):
                r2.window(frames, True)
            r2.drain()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            run.i = r2.i
            short["cfg2_window%d" % frames] = {
                "workload": "cfg2, %d-frame windows (%.2f ms of audio per a2cu_submit)" % (frames, frames / 48.0),
                "value": args.voices * frames * n2 / (sum(r2.dev_ms) / 1e3), "unit": "voice-samples/s",
                "ms_per_window": sum(r2.dev_ms) / n2,
                "e2e": {"value": args.voices * frames * n2 / (t1 - t0), "unit": "voice-samples/s",
                        "ms_per_window": 1e3 * (t1 - t0) / n2}}
    e.close()

    peak, peak_src = peaks()
    configs = {}
    if not args.no_configs and (fused or not multi):
        configs = secondary_configs(local, dev, world, rank, peak, args)
    configs.update(short)

    if rank == 0:
        cmd_bytes = 16 + 4                          # one event record + CSR offset per voice
        alg_bytes = args.voices * (2 * state_bytes + cmd_bytes) + WINDOW * 2 * 4
        achieved = alg_bytes / (k_ms / 1000.0) / 1e9
        traffic, traffic_src = ncu_traffic()
        cfg = config_of(args.voices)
        line = {
            "metric": "voice-samples/sec at 64-frame blocks",
            "value": value, "unit": "voice-samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": nwarm,
            "ms_per_step": dev_total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": cfg,
            "details": {
                "banks": "%d banks of %d voices, %.0f MB of voice state" % (nbanks, args.voices, state_mb),
                "warmup_windows": nwarm, "timed_windows": nwin, "ms_per_window": dev_total_ms / nwin,
                "timing": "value: CUDA-event spans of the kernels summed over the timed windows; e2e: wall time "
                          "of the timed region through write_all + a2cu_submit / a2cu_collect (2 windows in "
                          "flight, H2D + result into pinned host memory every window); max over ranks",
                "multi_gpu": ("voices sharded; root bus summed inside render_split's last CTA over NVLink peer "
                              "memory (a2cu_xchg_*), every rank receives the master block" if fused else
                              "voices sharded; NCCL int32 all-reduce of the root bus + separate root stage launch"
                              if multi else "single GPU"),
                "parity_check": parity,
            },
            "wall_ms_per_step": 1000.0 * (wall1 - wall0) / args.steps,
            "e2e": {"value": e2e, "unit": "voice-samples/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": host_total_ms / args.steps},
            "gpu_launches": int(launches),
            "kernel": "%s<%s>%s" % (kname, kernel_name,
                                     " (root stage%s fused)" % (" + NVLink exchange" if fused else "")
                                     if kname == "render_split" and (fused or not multi) else " + mix_root"),
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "traffic_source": "%s (ncu --set full, per launch; static: committed capture)" % traffic_src,
                "kernel": "%s<%s>" % (kname, kernel_name),
                "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "state is read and written once per 960-frame launch and the wavetable is staged in "
                        "shared memory: the kernel is bound by the filter12 recurrence latency and INT32 "
                        "issue, not by HBM (DESIGN.md 4/6, profiles/); the HBM-bound case is configs.gather",
            },
            "configs": configs,
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            frames = args.cpu_frames
            vs, sec, kind, used = run_reference_sample(args.voices, frames, 1)
            line["cpu_baseline"] = {
                "value": vs / sec, "unit": "voice-samples/s", "cores": 1, "kind": kind,
                "sample": "%d voices x %d frames of the same workload, one engine state "
                          "(single-threaded by design), a2_Run loop only" % (args.voices, frames)}
        print(json.dumps(line), flush=True)
    if multi:
        dist.barrier()
        dist.destroy_process_group()


# ---- library-collective baseline for A/B (--exchange nccl): cut path + NCCL + root stage ----
NCCL_STATE = {"dev_ms": [], "render_ms": [], "ring": None}


def nccl_window(e, run, dev, stream, reduce_root_bus, timed):
    import torch
    st = NCCL_STATE
    if st["ring"] is None:
        e.set_post_root_stage(False)
        st["ring"] = {
            "rootbus": torch.zeros((WINDOW, 2), dtype=torch.int32, device=dev),
            "master": torch.zeros((WINDOW, 2), dtype=torch.int32, device=dev),
            "host": torch.zeros((WINDOW, 2), dtype=torch.int32).pin_memory(),
            "a": torch.cuda.Event(enable_timing=True), "b": torch.cuda.Event(enable_timing=True)}
    g = st["ring"]
    L, n = e.L, len(run.banks)
    i = run.i
    cur = run.banks[i % n]
    L.a2cu_bank_enable(e.h, run.banks[(i - 1) % n], 0)
    L.a2cu_bank_enable(e.h, cur, 1)
    a = run.amp[(i // n) & 1]
    L.a2cu_bank_write_all(e.h, cur, 0, 2, a.ctypes.data, 0, L.a2cu_now(e.h), WINDOW << 8)
    t = e.submit_dev(WINDOW, BLOCK, g["rootbus"].data_ptr())
    g["a"].record(stream)
    reduce_root_bus(g["rootbus"])
    e.apply_root_stage(g["rootbus"].data_ptr(), g["master"].data_ptr(), WINDOW, BLOCK)
    g["b"].record(stream)
    g["host"].copy_(g["master"], non_blocking=True)
    stream.synchronize()
    e.collect_spans(t)
    if timed:
        st["dev_ms"].append(e.last_render_ms() + e.last_mix_ms() + g["a"].elapsed_time(g["b"]))
        st["render_ms"].append(e.last_render_ms())
    run.i += 1
    return g["host"].numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--voices", type=int, default=4096)
    ap.add_argument("--banks", type=int, default=0,
                    help="banks served round-robin, one per window (0: enough for 1.5x L2 of voice state)")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N > 1: in-kernel NVLink exchange (product) or the NCCL baseline")
    ap.add_argument("--gather-voices", type=int, default=131072)
    ap.add_argument("--cpu-frames", type=int, default=96000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary workloads")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
