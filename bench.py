#!/usr/bin/env python
"""bench.py - voice-samples/s of the voice-render hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path

Workload (config.workload "cfg2"): 4 096 voices wtosc -> filter12 -> panmix on
the shared 2 048-point saw, 48 kHz, 64-frame blocks, plus ONE control write
per voice per step (amplitude re-targeted and ramped across the step) so that
every step has real host->device input.  One step = one 960-frame window
(20 ms = 15 blocks of 64; one a2cu_submit) = voices x 960 voice-samples.

Numbers on the JSON line:
  value   voice-samples/s over the CUDA-event spans of the kernels (render +
          bus stage [+ NCCL reduce and root stage for N > 1]); the step's
          events are already in HBM when the span starts.  No L2 flush: every
          step renders a DIFFERENT bank of 4096 voices, round-robin over enough
          banks that their state exceeds the L2 by 1.5x (inputs larger than L2).
  e2e     the same metric through the public API with HOST buffers: per step
          a2cu_bank_write_all (host events), a2cu_submit (event staging, H2D,
          kernels, D2H of the int32 master block into pinned memory) and
          a2cu_collect (wait + copy to the caller's buffer), two windows in
          flight so the host stages step i+1 while the device renders step i;
          wall time of the whole timed region, max over ranks. (N > 1: the same
          pipeline with a2cu_submit_dev + NCCL all-reduce + root stage + D2H
          queued per step, three windows in flight.)
  roofline  HBM roofline of the dominant kernel (render_split<...>, one launch per step: the
          root stage is fused into its last CTA).
  cpu_baseline  the reference's own CPU render (oracle/_ref) on a bounded
          sample of the same workload, 1 core, rank 0, N = 1 only.
Multi-GPU (weak scaling): every rank renders its own bank; the raw stereo root
bus is summed with one NCCL all-reduce (int32) before the truncating root stage.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

STEP_MS = 20
RATE = 48000
STEP_FRAMES = STEP_MS * RATE // 1000     # 960 = 15 blocks of 64
BLOCK = 64
L2_BYTES = 126 * 1024 * 1024    # B200 L2


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per
    launch, from the committed `ncu --set full` summary of this workload."""
    path = os.path.join(ROOT, "profiles", "r01_v5_render_split_ncu.txt")
    try:
        tot = 0.0
        for ln in open(path):
            f = ln.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(f[2], 1)
                tot += float(f[1]) * mult
        return tot or None
    except Exception:
        return None


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
    from cases import bench_bank
    scn = bench_bank(nvoices, steps=steps, step_ms=STEP_MS, seed=seed)
    with open(path, "w") as f:
        f.write(scn.to_a2s())


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "a2render")
    return p if os.path.exists(p) else None


def run_reference_sample(nvoices, frames, shards, seed=324357):
    """Render `frames` frames of the cfg2 bench bank with the reference on
    `shards` host processes (one engine state is single-threaded, so the only
    legal parallelism is independent states on disjoint voice shards,
    audiality2.h.cmake:163-166). Returns (voice_samples, seconds, kind)."""
    exe = ref_binary()
    steps = (frames + STEP_FRAMES - 1) // STEP_FRAMES
    if exe is None:
        # plain-C port (single thread)
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
    frames = RATE          # one second of audio per step (bounded sample)
    times = []
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
        "config": {"workload": "cfg2: %d voices wtosc->filter12->panmix, saw 2048-pt, 48 kHz, "
                               "64-frame blocks, 1 amplitude write/voice/20 ms" % args.voices},
        "cpu_baseline": {"value": value, "unit": "voice-samples/s", "cores": used,
                         "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "voice-samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def bench_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from audiality2_b200 import engine as eng
    from audiality2_b200.workloads import cfg2_bank
    from audiality2_b200.parallel import reduce_root_bus
    from scenarios import autowire

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    e = eng.Engine(RATE, 2, device=local)
    stream = torch.cuda.current_stream()
    e.set_stream(stream.cuda_stream)
    e.set_timing(True)
    # Round-robin over R independent banks of `voices` voices (one per step), so
    # that the per-voice state a step reads was last touched R steps ago and the
    # state of all banks together (R x voices x state bytes) exceeds the 126 MB
    # L2 by 1.5x: "inputs larger than L2" instead of an L2 flush between steps.
    b = cfg2_bank(args.voices, seed=324357 + rank)
    w = e.builtin_wave(b["wave"])
    chain = autowire(list(b["kinds"]))
    banks = []
    nbanks = args.banks
    while True:
        r = len(banks)
        bb = b if r == 0 else cfg2_bank(args.voices, seed=324357 + rank + 1000 * r) if r < 8 else b
        bank = e.new_bank(chain, args.voices)
        e.write_all(bank, 0, 0, [w << 16], dur=STEP_FRAMES << 8)
        e.write_all(bank, 0, 1, bb["pitch"])
        e.write_all(bank, 0, 2, [bb["amp"]])
        e.write_all(bank, 1, 0, bb["cutoff"])
        e.write_all(bank, 1, 1, [bb["q"]])
        e.write_all(bank, 2, 1, bb["pan"])
        e.bank_enable(bank, False)
        banks.append(bank)
        if nbanks <= 0:
            nbanks = int(1.5 * L2_BYTES / (e.bank_state_bytes(bank) * args.voices)) + 1
        if len(banks) >= nbanks:
            break
    bank = banks[0]
    state_mb = nbanks * e.bank_state_bytes(bank) * args.voices / 1e6
    multi = world > 1
    if multi:
        e.set_post_root_stage(False)
    RING = 3                # multi-GPU: windows in flight (device buffers + events per slot)
    rootbus = [torch.zeros((STEP_FRAMES, 2), dtype=torch.int32, device=dev) for _ in range(RING)]
    master = [torch.zeros((STEP_FRAMES, 2), dtype=torch.int32, device=dev) for _ in range(RING)]
    host_out = [torch.zeros((STEP_FRAMES, 2), dtype=torch.int32).pin_memory() for _ in range(RING)]
    amp = [np.array([b["amp"] // 2], dtype=np.int32), np.array([b["amp"]], dtype=np.int32)]
    L = e.L
    ev_a = [torch.cuda.Event(enable_timing=True) for _ in range(RING)]
    ev_b = [torch.cuda.Event(enable_timing=True) for _ in range(RING)]
    fin = [torch.cuda.Event() for _ in range(RING)]
    dev_ms, host_s, render_ms = [], [], []
    state = {"step": 0}

    out_host = np.empty((STEP_FRAMES, 2), dtype=np.int32)
    pending = []            # tickets in flight (single GPU: pipelined a2cu_submit / a2cu_collect)

    def collect_one(timed):
        e.collect(pending.pop(0), out_host)        # waits for THAT window's D2H only
        if timed:
            dev_ms.append(e.last_render_ms() + e.last_mix_ms())
            render_ms.append(e.last_render_ms())

    def one_step(timed):
        i = state["step"]
        cur = banks[i % nbanks]
        # this step's bank: resume it, pause the one of the previous step
        L.a2cu_bank_enable(e.h, banks[(i - 1) % nbanks], 0)
        L.a2cu_bank_enable(e.h, cur, 1)
        a = amp[(i // nbanks) & 1]
        if not multi:
            # --- the step, through the public C ABI, host buffers in and out ---
            L.a2cu_bank_write_all(e.h, cur, 0, 2, a.ctypes.data, 0, L.a2cu_now(e.h), STEP_FRAMES << 8)
            # event staging + H2D + kernels + D2H queued; the host goes on to
            # prepare the next step while the device renders this one
            pending.append(e.submit(STEP_FRAMES, BLOCK))
            if len(pending) > 2:
                collect_one(timed)
        else:
            # N > 1: the same pipeline with the exchange step on the stream: raw stereo root bus of
            # this rank's voices -> NCCL int32 all-reduce over NVLink -> truncating root stage -> D2H
            k = i % RING
            if inflight[k] is not None:
                finish_slot(k)
            L.a2cu_bank_write_all(e.h, cur, 0, 2, a.ctypes.data, 0, L.a2cu_now(e.h), STEP_FRAMES << 8)
            t = e.submit_dev(STEP_FRAMES, BLOCK, rootbus[k].data_ptr())
            ev_a[k].record(stream)
            reduce_root_bus(rootbus[k])
            e.apply_root_stage(rootbus[k].data_ptr(), master[k].data_ptr(), STEP_FRAMES, BLOCK)
            ev_b[k].record(stream)
            host_out[k].copy_(master[k], non_blocking=True)
            fin[k].record(stream)
            inflight[k] = (t, timed)
        state["step"] += 1

    inflight = [None] * RING

    def finish_slot(k):
        t, timed = inflight[k]
        inflight[k] = None
        fin[k].synchronize()                        # this window's result is in host_out[k]
        e.collect_spans(t)
        if timed:
            dev_ms.append(e.last_render_ms() + e.last_mix_ms() + ev_a[k].elapsed_time(ev_b[k]))
            render_ms.append(e.last_render_ms())

    def drain(timed):
        while pending:
            collect_one(timed)
        base = state["step"]
        for j in range(RING):                       # oldest first
            k = (base + j) % RING
            if inflight[k] is not None:
                finish_slot(k)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # >= max(3, W) warm-up steps, then chunks of 50 more until the GPU has been
    # under this load for ~1.5 s, so nvidia-smi (100 ms period) samples clocks
    # under load. Rank 0 decides and broadcasts (all ranks must run the same
    # number of collective steps).
    nwarm = max(3, args.warmup, nbanks)     # every bank is rendered once before timing
    t_w = time.perf_counter()
    for _ in range(nwarm):
        one_step(False)
    drain(False)
    while True:
        go = torch.tensor([1 if time.perf_counter() - t_w < 1.5 and nwarm < 100000 else 0], device=dev)
        if multi:
            dist.broadcast(go, 0)
        if not int(go.item()):
            break
        for _ in range(50):
            one_step(False)
        drain(False)
        nwarm += 50
    if multi:
        dist.barrier()
    torch.cuda.synchronize()
    l0, h0, d0 = e.launches, e.h2d_bytes, e.d2h_bytes
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        one_step(True)
    drain(True)
    torch.cuda.synchronize()
    if multi:
        dist.barrier()
    wall1 = time.perf_counter()
    host_s.append(wall1 - wall0)            # pipelined: the whole timed region is the e2e time
    launches = e.launches - l0
    h2d = (e.h2d_bytes - h0) / args.steps
    d2h = (e.d2h_bytes - d0) / args.steps
    if multi:
        d2h = STEP_FRAMES * 2 * 4
    clk = clocks.stop() if rank == 0 else None

    tot = torch.tensor([sum(dev_ms), sum(host_s) * 1000.0], dtype=torch.float64, device=dev)
    if multi:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    dev_total_ms, host_total_ms = [float(x) for x in tot.tolist()]
    vs_total = float(args.voices) * STEP_FRAMES * args.steps * world
    value = vs_total / (dev_total_ms / 1000.0)
    e2e = vs_total / (host_total_ms / 1000.0)

    if rank == 0:
        peak, peak_src = peaks()
        kname = "render_split" if e.split_launches else "render_bank"
        state_bytes = e.bank_state_bytes(bank)
        cmd_bytes = 16 + 4                          # one event record + CSR offset per voice
        alg_bytes = args.voices * (2 * state_bytes + cmd_bytes) + STEP_FRAMES * 2 * 4
        k_ms = statistics.mean(render_ms)
        achieved = alg_bytes / (k_ms / 1000.0) / 1e9
        line = {
            "metric": "voice-samples/sec at 64-frame blocks",
            "value": value, "unit": "voice-samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": nwarm,
            "ms_per_step": dev_total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {
                "workload": "cfg2: %d voices/GPU wtosc->filter12->panmix, saw 2048-pt, 48 kHz, "
                            "64-frame blocks, 1 amplitude write/voice/20 ms" % args.voices,
                "step": "%d frames (15 blocks of 64) per window (one a2cu_submit)" % STEP_FRAMES,
                "l2": "no flush: inputs larger than L2 - every step renders a different bank of %d voices, "
                      "round-robin over %d banks whose per-voice state totals %.0f MB (1.5x the 126 MB L2)"
                      % (args.voices, nbanks, state_mb),
                "timing": "value: CUDA-event spans of kernels summed over steps; e2e: wall time of the "
                          "timed region through write_all + a2cu_submit/a2cu_collect (2 windows in flight, "
                          "H2D + D2H every step); max over ranks",
                "multi_gpu": "voices sharded, one NCCL int32 all-reduce of the root bus per step"
                             if multi else "single GPU",
            },
            "wall_ms_per_step": 1000.0 * (wall1 - wall0) / args.steps,
            "e2e": {"value": e2e, "unit": "voice-samples/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": host_total_ms / args.steps},
            "gpu_launches": int(launches),
            "kernel": "%s<%s>%s" % (kname, e.bank_kernel_name(bank),
                                     " (root stage fused)" if not multi and kname == "render_split" else " + mix_root"),
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                "traffic_source": "profiles/r01_v5_render_split_ncu.txt (ncu --set full, per launch)",
                "kernel": "%s<%s>" % (kname, e.bank_kernel_name(bank)),
                "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "state is read and written once per 960-frame launch and the wavetable "
                        "is L1/L2 resident: the kernel is bound by the filter12 recurrence "
                        "latency and INT32 issue, not by HBM (DESIGN.md 4/6, profiles/)",
            },
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
    e.close()
    if multi:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--voices", type=int, default=4096)
    ap.add_argument("--banks", type=int, default=0,
                    help="banks served round-robin, one per step (0: enough for 1.5x L2 of voice state)")
    ap.add_argument("--cpu-frames", type=int, default=96000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
