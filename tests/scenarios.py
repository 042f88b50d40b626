"""Scenario model shared by the parity tests.

A Scenario describes a bank of voices (unit chains), the control writes each
voice's A2S program would perform and when, and the render settings.  It can be
turned into

  * an .a2s script for the reference (`to_a2s`),  run through oracle/_ref/a2render
  * an event list for the C port              (`run_oracle`)
  * calls into the CUDA engine's C ABI         (`run_cuda`, tests/ only on GPU)

so the three can be compared bit for bit on identical inputs.  All values are
16:16 fixed point integers, exactly what the VM registers would hold; script
literals are printed so that the compiler's floor(v * 65536 + .5)
(compiler.c:494-503) reproduces them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from audiality2_b200.chains import (GENERATORS, FM_OPS, UNIT_IO, REGS, KIND_CODE,  # noqa: E402,F401
                                     autowire)


ROOT_WAKE_PERIOD = 1000000     # core.c:1195


def fx(x):
    """float -> 16:16 like a2c_Num2VM (compiler.c:494-503)."""
    return int(np.floor(x * 65536.0 + 0.5))


def lit(v):
    """16:16 int -> script literal that compiles back to exactly v."""
    s = "%.9f" % (v / 65536.0)   # a2_GetNum overflows at 10 decimals (compiler.c:876,889)
    s = s.rstrip("0").rstrip(".")
    return s if s not in ("", "-0") else "0"


class Voice:
    """One leaf voice: a struct (list of unit kinds) and a program: a list of
    steps. Steps:
        ("set", unit, reg, value)         '@reg value'  (immediate, dur 0)
        ("ramp", unit, reg, value)        'reg value'   (applied at next delay)
        ("d", ms_16_16)                   'd ms'
    The program ends with an endless sleep."""

    def __init__(self, kinds, steps, group=-1, transpose=0):
        self.kinds = list(kinds)
        self.steps = list(steps)
        self.group = group
        self.transpose = transpose


class Scenario:
    def __init__(self, samplerate=48000, channels=2, buffer=64, frames=4800,
                 noiseseed=None):
        self.samplerate = samplerate
        self.channels = channels
        self.buffer = buffer
        self.frames = frames
        self.noiseseed = noiseseed
        self.transpose = 0        # Song-level 'tr', inherited by every voice
        self.voices = []
        self.ngroups = 0
        self.group_steps = {}     # group -> steps on its panmix (unit 0)
        self.group_fbd = {}       # group -> 7 fbdelay registers (16:16) of a song-level chain
        self.waves = []           # names of builtin waves, index = wave id
        self.uploaded = None      # (type, period, flags, length, seed) of the one sampled wave
        # writes to the ROOT driver's panmix, (time 24:8, reg, value, dur): only an API client can
        # do that (a2_Send to the root voice), so these exist for the port and the CUDA engine only
        self.root_writes = []

    def upload(self, wtype, period, flags, length, seed):
        """One pseudo-random sampled wave, uploaded through a2_UploadWave by the reference
        harness (`a2render -U`), a2cu_wave_upload / a2o_upload_wave here. Returns its wave id."""
        assert self.uploaded is None
        self.uploaded = (wtype, period, flags, length, seed)
        self.waves.append(self.uploaded)
        return len(self.waves) - 1

    def uploaded_data(self):
        """The harness's generator (oracle/a2render.c, -U): LCG, int16 in [-25000, 25000]."""
        _, _, _, length, seed = self.uploaded
        out = np.empty(length, dtype=np.int16)
        x = seed & 0xffffffff
        for k in range(length):
            x = (x * 1664525 + 1013904223) & 0xffffffff
            out[k] = (x >> 16) % 50001 - 25000
        return out

    def wave(self, name):
        if name not in self.waves:
            self.waves.append(name)
        return self.waves.index(name)

    def add_group(self, steps=(), fbdelay=None):
        """fbdelay: (fbdelay, ldelay, rdelay [ms], drygain, fbgain, lgain, rgain), 16:16 - makes the
        group the song-level chain { inline 0 *; fbdelay * *; panmix * > } (k2trance.a2s:920-924)."""
        self.group_steps[self.ngroups] = list(steps)
        if fbdelay is not None:
            self.group_fbd[self.ngroups] = [int(x) for x in fbdelay]
        self.ngroups += 1
        return self.ngroups - 1

    def add_voice(self, kinds, steps, group=-1):
        self.voices.append(Voice(kinds, steps, group, self.transpose))
        return len(self.voices) - 1

    # ------------------------------------------------------------------
    def msdur(self):
        return int(np.float32(self.samplerate) * np.float32(65.536) + np.float32(.5))

    def ms2t(self, d):
        """core.c:1127-1130"""
        return ((d * self.msdur() + 0x7fffff) >> 24) & 0xffffffff

    def _unit_names(self, kinds):
        names, seen = [], {}
        for k in kinds:
            c = seen.get(k, 0)
            seen[k] = c + 1
            names.append(None if c == 0 else "u%d" % (len(names) + 1))
        return names

    def _regname(self, kinds, names, unit, reg):
        r = REGS[kinds[unit]][reg]
        return r if names[unit] is None else "%s.%s" % (names[unit], r)

    def _steps_to_events(self, steps, target, kind_write, kinds):
        """Mirror of the VM: returns event tuples
        (time, kind, target, unit, reg, value, dur)."""
        ev = []
        t = 0
        pending = []

        def flat(steps):
            for st in steps:
                if st[0] == "loop":
                    for _ in range(st[1]):
                        for x in flat(st[2]):
                            yield x
                else:
                    yield st

        for st in flat(steps):
            if st[0] == "set":
                _, u, r, v = st
                pending = [p for p in pending if (p[0], p[1]) != (u, r)]
                ev.append((t, kind_write, target, u, r, v, 0))
            elif st[0] == "ramp":
                _, u, r, v = st
                for p in pending:
                    if (p[0], p[1]) == (u, r):
                        p[2] = v
                        break
                else:
                    pending.append([u, r, v])
            elif st[0] == "d":
                dt = self.ms2t(st[1])
                for u, r, v in pending:
                    ev.append((t, kind_write, target, u, r, v, dt))
                pending = []
                t += dt
                ev.append((t, 1, target, 0, 0, 0, 0))   # wake-up
        return ev

    def events(self):
        from oracle import a2oracle as ao
        ev = []
        for vi, v in enumerate(self.voices):
            ev += self._steps_to_events(v.steps, vi, ao.EV_WRITE, v.kinds)
        for g, steps in self.group_steps.items():
            e = self._steps_to_events(steps, g, ao.EV_GROUPWRITE, ["panmix"])
            # the group's own wake-ups: GROUPWRITE with reg -1 (split only);
            # its children are split implicitly by the inline recursion
            # (core.c:1769-1776)
            for x in e:
                if x[1] == ao.EV_WAKE:
                    ev.append((x[0], ao.EV_GROUPWRITE, g, 0, -1, 0, 0))
                else:
                    ev.append(x)
        # The root voice runs a2_rootdriver, whose program falls into OP_END
        # while attached: it re-arms its wake-up every 1000000 (24:8) time
        # units (core.c:1191-1216), splitting every voice's segments there.
        for (tw, reg, value, dur) in self.root_writes:
            ev.append((tw, ao.EV_ROOTWRITE, 0, 0, reg, value, dur))
        t = ROOT_WAKE_PERIOD
        while t < (self.frames << 8):
            ev.append((t, ao.EV_ROOTWRITE, 0, 0, -1, 0, 0))
            t += ROOT_WAKE_PERIOD
        order = sorted(range(len(ev)), key=lambda i: ev[i][0])
        arr = np.zeros(len(ev), dtype=ao.EVENT_DTYPE)
        for j, i in enumerate(order):
            arr[j] = ev[i]
        return arr

    # ------------------------------------------------------------------
    def to_a2s(self):
        """Script whose Song() spawns every group and voice at time 0."""
        L = ['def title "generated"', 'def a2sversion "1.9"', ""]
        # an uploaded wave is known to the script only as a handle: program argument UW
        arg = "UW" if self.uploaded is not None else ""

        def wname(idx):
            w = self.waves[idx]
            return "UW" if isinstance(w, tuple) else w

        def body(kinds, steps, names, indent="\t", final=True):
            out = []
            for st in steps:
                if st[0] == "loop":
                    out.append("%s%d {" % (indent, st[1]))
                    out += body(kinds, st[2], names, indent + "\t", False)
                    out.append("%s}" % indent)
                elif st[0] == "set":
                    rn = self._regname(kinds, names, st[1], st[2])
                    val = wname(st[3] >> 16) if st[2] == 0 and \
                        kinds[st[1]] == "wtosc" else lit(st[3])
                    out.append("%s@%s %s" % (indent, rn, val))
                elif st[0] == "ramp":
                    rn = self._regname(kinds, names, st[1], st[2])
                    val = wname(st[3] >> 16) if st[2] == 0 and \
                        kinds[st[1]] == "wtosc" else lit(st[3])
                    out.append("%s%s %s" % (indent, rn, val))
                elif st[0] == "d":
                    out.append("%sd %s" % (indent, lit(st[1])))
            if final:
                out.append("\tfor { d 30000 }")
            return out

        for vi, v in enumerate(self.voices):
            names = self._unit_names(v.kinds)
            units = "; ".join(k if nm is None else "%s %s" % (k, nm)
                              for k, nm in zip(v.kinds, names))
            L.append("V%d(%s)" % (vi, arg))
            L.append("{")
            L.append("\tstruct { %s }" % units)
            L += body(v.kinds, v.steps, names)
            L.append("}")
            L.append("")
        # Groups: { inline 0 *; panmix * > } == a2_groupdriver numerically
        for g in range(self.ngroups):
            members = [vi for vi, v in enumerate(self.voices) if v.group == g]
            L.append("G%d(%s)" % (g, arg))
            L.append("{")
            if g in self.group_fbd:
                L.append("\tstruct { inline 0 *; fbdelay D * *; panmix * > }")
                for rn, val in zip(("fbdelay", "ldelay", "rdelay", "drygain", "fbgain", "lgain",
                                    "rgain"), self.group_fbd[g]):
                    L.append("\tD.%s %s" % (rn, lit(val)))
                L.append("\tset")
            else:
                L.append("\tstruct { inline 0 *; panmix * > }")
            for vi in members:
                L.append("\tV%d %s" % (vi, arg))
            L += body(["panmix"], self.group_steps[g], [None])
            L.append("}")
            L.append("")
        # Spawn order: my engines process newest-first at every level, like
        # a2_VoiceNew's head insertion; creation order here = index order.
        L.append("export Song(%s)" % arg)
        L.append("{")
        if self.transpose:
            L.append("\ttr %s" % lit(self.transpose))
        # root children in creation order: groups are created before the
        # voices that reference them, voices with group -1 in index order
        items = []
        for g in range(self.ngroups):
            items.append(("g", g))
        for vi, v in enumerate(self.voices):
            if v.group < 0:
                items.append(("v", vi))
        if len(items) <= 100:
            for kind, idx in items:
                L.append("\t%s%d %s" % ("G" if kind == "g" else "V", idx, arg))
        else:
            # One program may run at most A2_INSLIMIT = 1000 VM instructions
            # between timing points (config.h:119): spawn through helper
            # voices without units (their children inherit the root bus).
            nsp = (len(items) + 63) // 64
            for k in range(nsp):
                L.append("\tS%d %s" % (k, arg))
        L.append("\tfor { d 30000 }")
        L.append("}")
        if len(items) > 100:
            head = L[:3]
            sp = []
            for k in range(nsp):
                sp.append("S%d(%s)" % (k, arg))
                sp.append("{")
                for kind, idx in items[k * 64:(k + 1) * 64]:
                    sp.append("\t%s%d %s" % ("G" if kind == "g" else "V", idx, arg))
                sp.append("\tfor { d 30000 }")
                sp.append("}")
                sp.append("")
            # programs must be defined before use: voices, groups, spawners, Song
            song_at = L.index("export Song()")
            L = L[:song_at] + sp + L[song_at:]
        return "\n".join(L) + "\n"


def run_ref(scn, path):
    from oracle import a2oracle as ao
    with open(path, "w") as f:
        f.write(scn.to_a2s())
    out, info = ao.ref_render(path, "Song", samplerate=scn.samplerate,
                              channels=scn.channels, buffer=scn.buffer,
                              frames=scn.frames, noiseseed=scn.noiseseed,
                              upload=scn.uploaded)
    if info["rt_error"]:
        raise RuntimeError("reference reported RT error %d" % info["rt_error"])
    return out


def build_engine(scn, eng, kindmod):
    """Create waves, groups and voices of `scn` in `eng` (oracle or cuda
    wrapper: both expose builtin_wave/new_group/new_voice)."""
    for w in scn.waves:
        if isinstance(w, tuple):
            eng.upload_wave(w[0], w[1], w[2], scn.uploaded_data())
        else:
            eng.builtin_wave(w)
    for g in range(scn.ngroups):
        eng.new_group()
        if g in scn.group_fbd:
            eng.group_fbdelay(g, scn.group_fbd[g])
    for v in scn.voices:
        eng.new_voice(autowire(v.kinds), transpose=v.transpose, substart=0,
                      group=v.group)


def run_oracle(scn):
    from oracle import a2oracle as ao
    o = ao.Oracle(scn.samplerate, scn.channels)
    if scn.noiseseed is not None:
        o.set_noiseseed(scn.noiseseed)
    build_engine(scn, o, ao)
    out = o.render(scn.events(), scn.frames, scn.buffer)
    o.close()
    return out


def run_cuda(scn, window=None, split=True, stats=None, pipelined=False):
    """Render `scn` with the CUDA engine through its C ABI (ctypes).
    window: frames per a2cu_run call (multiple of scn.buffer), default all.
    split=False forces the one-thread-per-voice kernel (render_bank).
    pipelined: windows go through a2cu_submit / a2cu_collect, up to 3 in flight."""
    from audiality2_b200 import engine as eng
    from oracle import a2oracle as ao
    e = eng.Engine(scn.samplerate, scn.channels)
    try:
        if not split:
            e.set_split(False)
        if scn.noiseseed is not None:
            e.set_noiseseed(scn.noiseseed)
        for w in scn.waves:
            if isinstance(w, tuple):
                e.upload_wave(w[0], w[1], w[2], scn.uploaded_data())
            else:
                e.builtin_wave(w)
        if scn.group_fbd:
            raise NotImplementedError("bank mode has no group-level fbdelay (drop-in mode only)")
        for _ in range(scn.ngroups):
            e.new_group()
        # one bank per distinct voice structure, slots in voice order
        chains, where = {}, []
        for v in scn.voices:
            key = tuple(v.kinds)
            chains.setdefault(key, []).append(v)
            where.append((key, len(chains[key]) - 1))
        bank_of = {}
        for key, vs in chains.items():
            bank_of[key] = e.new_bank(autowire(list(key)), len(vs),
                                      transpose=[v.transpose for v in vs],
                                      group=[v.group for v in vs])
        for ev in scn.events():
            t, kind, tgt = int(ev["time"]), int(ev["kind"]), int(ev["voice"])
            if kind == ao.EV_WRITE:
                key, slot = where[tgt]
                e.write(bank_of[key], slot, int(ev["unit"]), int(ev["reg"]),
                        int(ev["value"]), t, int(ev["dur"]))
            elif kind == ao.EV_WAKE:
                key, slot = where[tgt]
                e.wake(bank_of[key], slot, t)
            elif kind == ao.EV_GROUPWRITE:
                e.group_write(tgt, int(ev["reg"]), int(ev["value"]), t, int(ev["dur"]))
            elif kind == ao.EV_ROOTWRITE and int(ev["reg"]) >= 0:
                e.root_write(int(ev["reg"]), int(ev["value"]), t, int(ev["dur"]))
            # root wake-ups (reg < 0) are the engine's own root_wake_period
        if window is None:
            out = e.run(scn.frames, scn.buffer)
        elif pipelined:
            parts, tickets, done = [], [], 0
            while done < scn.frames:
                n = min(window, scn.frames - done)
                tickets.append(e.submit(n, scn.buffer))
                done += n
                if len(tickets) > 2:
                    parts.append(e.collect(tickets.pop(0)))
            while tickets:
                parts.append(e.collect(tickets.pop(0)))
            out = np.concatenate(parts, axis=0)
        else:
            parts, done = [], 0
            while done < scn.frames:
                n = min(window, scn.frames - done)
                parts.append(e.run(n, scn.buffer))
                done += n
            out = np.concatenate(parts, axis=0)
        if stats is not None:
            stats["launches"] = e.launches
            stats["split_launches"] = e.split_launches
        return out
    finally:
        e.close()


def _build_cuda_shard(scn, e, voice_ids):
    """Create the waves and the voices `voice_ids` of `scn` (root-level voices only) in engine
    `e` and queue their events. Returns nothing; mirrors run_cuda's setup."""
    from oracle import a2oracle as ao
    for w in scn.waves:
        if isinstance(w, tuple):
            e.upload_wave(w[0], w[1], w[2], scn.uploaded_data())
        else:
            e.builtin_wave(w)
    assert scn.ngroups == 0, "sharding cuts at the root bus: root-level voices only"
    chains, where = {}, {}
    for vi in voice_ids:
        v = scn.voices[vi]
        key = tuple(v.kinds)
        chains.setdefault(key, []).append(v)
        where[vi] = (key, len(chains[key]) - 1)
    bank_of = {}
    for key, vs in chains.items():
        bank_of[key] = e.new_bank(autowire(list(key)), len(vs),
                                  transpose=[v.transpose for v in vs])
    for ev in scn.events():
        t, kind, tgt = int(ev["time"]), int(ev["kind"]), int(ev["voice"])
        if kind == ao.EV_WRITE and tgt in where:
            key, slot = where[tgt]
            e.write(bank_of[key], slot, int(ev["unit"]), int(ev["reg"]), int(ev["value"]), t,
                    int(ev["dur"]))
        elif kind == ao.EV_WAKE and tgt in where:
            key, slot = where[tgt]
            e.wake(bank_of[key], slot, t)
        elif kind == ao.EV_ROOTWRITE and int(ev["reg"]) >= 0:
            # every shard sees the root's writes: they cut all voices' segments, and the
            # root panmix runs (identically) on every rank after the exchange
            e.root_write(int(ev["reg"]), int(ev["value"]), t, int(ev["dur"]))


def run_cuda_sharded(scn, nshards=2, window=None, split=True, mode="fused", stats=None, pipelined=False):
    """Render `scn` on `nshards` engines of ONE process (all on cuda:0, one CUDA stream each),
    voices dealt round-robin, and return every shard's master output.
      mode "fused": in-kernel exchange over peer memory (a2cu_xchg_*): each window is submitted on
                    all engines before any is collected - the kernels wait for each other.
      mode "cut":   set_post_root_stage(0) per shard, host-side integer sum of the raw root buses,
                    a2cu_apply_root_stage on shard 0 (the library-collective baseline's data path)."""
    import torch
    from audiality2_b200 import engine as eng
    window = window or scn.frames
    engines, streams = [], []
    try:
        for s in range(nshards):
            e = eng.Engine(scn.samplerate, scn.channels)
            st = torch.cuda.Stream()
            e.set_stream(st.cuda_stream)
            if not split:
                e.set_split(False)
            _build_cuda_shard(scn, e, range(s, len(scn.voices), nshards))
            engines.append(e)
            streams.append(st)
        if mode == "fused":
            for s, e in enumerate(engines):
                e.xchg_create(s, nshards, window, timeout_ms=5000)
            for e in engines:
                e.xchg_connect_local(engines)
            outs = [[] for _ in engines]
            done = 0
            inflight = [[] for _ in engines]
            while done < scn.frames:
                n = min(window, scn.frames - done)
                for s, e in enumerate(engines):
                    inflight[s].append(e.submit(n, scn.buffer))
                done += n
                # pipelined: up to three windows in flight per engine - a window's root stage then runs
                # at the tail of the NEXT window's kernel (lagged exchange); otherwise every window is
                # collected at once (a2cu_collect finishes it with the drain kernel)
                while inflight[0] and (not pipelined or len(inflight[0]) > 2):
                    for s, e in enumerate(engines):
                        outs[s].append(e.collect(inflight[s].pop(0)))
            while inflight[0]:
                for s, e in enumerate(engines):
                    outs[s].append(e.collect(inflight[s].pop(0)))
            if stats is not None:
                stats["launches"] = [e.launches for e in engines]
                stats["split_launches"] = [e.split_launches for e in engines]
            return [np.concatenate(o, axis=0) for o in outs]
        assert mode == "cut"
        for e in engines:
            e.set_post_root_stage(False)
        parts, done = [], 0
        dev = torch.device("cuda", 0)
        while done < scn.frames:
            n = min(window, scn.frames - done)
            raw = [e.run(n, scn.buffer).astype(np.int64) for e in engines]
            total = sum(raw)
            total = ((total + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int32)   # int32 wrap-around sum
            bus = torch.from_numpy(total).to(dev)
            master = torch.zeros((n, scn.channels), dtype=torch.int32, device=dev)
            with torch.cuda.stream(streams[0]):
                engines[0].apply_root_stage(bus.data_ptr(), master.data_ptr(), n, scn.buffer)
            streams[0].synchronize()
            parts.append(master.cpu().numpy())
            done += n
        return [np.concatenate(parts, axis=0)]
    finally:
        for e in engines:
            e.close()
