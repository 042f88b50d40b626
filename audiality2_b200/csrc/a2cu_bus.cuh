// a2cu_bus.cuh - bus-stage kernels (non-template): group / root panmix over a window, the
// sharded root stage, and the drop-in mode bus interpreter. Included by a2cu_engine.cu only.
#pragma once
#include "a2cu_kernels.cuh"

namespace a2cu {

// groups: { inline 0 *; panmix * *; xinsert * > } -> add into the root bus.
// grid = ngroups
__global__ void __launch_bounds__(256) mix_groups(const MixParams P) {
    const int g = blockIdx.x;
    int *root = P.acc;
    int *in = P.acc + (size_t)(1 + g) * P.W * 2;
    const bool clear = P.clear != 0;
    pm_bus(P, g, P.gstate + g * 8, in, false, [&](int f, int r0, int r1) {
        atomicAdd(root + f * 2, r0);
        atomicAdd(root + f * 2 + 1, r1);
        if (clear) { in[f * 2] = 0; in[f * 2 + 1] = 0; }
    });
}

// root: { inline 0 *|2; panmix * *|2 1; xinsert * > } into the cleared master;
// with root_stage == 0 the raw root bus is copied out (multi-GPU cut).
__global__ void __launch_bounds__(256) mix_root(const MixParams P) {
    root_stage(P, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, blockIdx.x == 0);
}

// Self-test hook (a2cu_debug_f12_coeff): the filter12 coefficient exactly as the kernels obtain it.
__global__ void f12_coeff_probe(Ctx c, const int *cutoff_values, int n, int *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f12_coeff(c, cutoff_values[i]);
}

// Root stage of a sharded render outside render_split's fused tail: ONE CTA sums the root
// bus over all ranks through peer memory (xchg_root_bus), then runs the root panmix.
__global__ void __launch_bounds__(512) mix_root_xchg(const MixParams P, const XchgParams X) {
    xchg_root_bus(X, P.acc, P.W, threadIdx.x, blockDim.x);
    root_stage(P, threadIdx.x, blockDim.x, true);
}

// The last window of a lagged sequence (nothing follows that could finish it).
__global__ void __launch_bounds__(512) xchg_drain(const XchgParams X, int *rstate, int channels, int root_stage_on,
                                                  int out_fmt) {
    xchg_finish_previous(X, rstate, channels, root_stage_on, threadIdx.x, blockDim.x, out_fmt);
}

// ---------------------------------------------------------------------------
// Drop-in mode bus stage: the bus-level Process()/write calls the host made
// during its tree walk.
//
// Commands are grouped into RUNS: all commands one voice received in this
// flush, in host order. Data only flows upwards in the voice tree (a voice's
// output is `+=`-ed into its parent's bus, core.c:1763-1776), so runs of the
// same nest level are independent of each other: the engine launches
// bus_level once per nest level, deepest first, one CTA per run. Adds into a
// bus another run of the level may also add to (wire-outs into the parent's
// bus) are integer atomics - order-free, bit-exact.
//
//   BUS_PM_*   bus-level panmix (panmix.c, all variants): control part on
//              thread 0, frames in closed form across the CTA
//   BUS_U_*    any other replaced unit called outside a fused leaf voice
//              ({inline; filter12}, {inline; wtosc; panmix}, {wtosc; panmix;
//              dcblock}, ...): one Process() call of ONE unit, exactly as the
//              reference runs it (unit by unit over the segment,
//              core.c:1875-1876), with the voice's scratch channels held in a
//              device bus row instead of st->scratch[nest]. The unit templates
//              are the code the fused kernels use, instantiated in replace
//              mode; add / wire-out (A2_PROCADD, A2_IO_WIREOUT) are applied
//              here. Recurrences run on thread 0.
//              kind A2CU_FBDELAY: units/fbdelay.c:68-127; frames run in
//              parallel when no tap of the call can see a sample written by
//              the same call, else on thread 0.
//   BUS_ADD    dst bus += src bus (an adding `inline` after device scratch,
//              host contributions uploaded into a staging row)
// ---------------------------------------------------------------------------
enum BusOp { BUS_PM_PROC = 0, BUS_PM_WRITE = 1, BUS_U_INIT = 2, BUS_U_WRITE = 3, BUS_U_SEED = 4, BUS_U_RUN = 5,
             BUS_ADD = 6 };
struct BusCmd {
    int op, pm;             // pm: index of the panmix instance state / generic unit state
    int nin, nout, add;     // add: bit 0 A2_PROCADD, bit 1 wire-out (BUS_U_RUN)
    int in_bus, out_bus;    // device bus indices (stereo rows of acc); BUS_U_RUN: scratch bus, wire target
    int frame, frames;
    int reg, value, start, dur;
    int kind;               // BUS_U_*: unit kind (A2CU_*)
    int run;                // host side: run this command belongs to
    int pad;
};
struct BusRun { unsigned begin, count; };

constexpr int kUnitWords = 64;      // state words reserved per generic unit (fm4: 16 x 4)
constexpr int kFbdKind = 5;         // A2CU_FBDELAY
constexpr int kFbdSize = 131072;    // A2FBD_BUFSIZE, fbdelay.c:26

struct BusVmParams {
    const BusCmd *cmds;
    const BusRun *runs;     // this level's runs; grid = number of runs
    int *acc;               // [bus][64][2]
    int *pmstate;           // [pm][8]
    int *ustate;            // [unit][kUnitWords]
    Ctx ctx;                // fmsine points at the global table here
};

// One call of one unit - Initialize / write / Process - on a voice's scratch channels.
//   sp, w0   the unit's state: word w0 of the voice (AoS unit block: stride 1, w0 = 0; bank SoA: the bank's stride)
//   scr      the voice's two scratch channels of this fragment, [frame][2]
//   out      the voice's output bus of this fragment, [frame][2] (wire-out units add with atomics)
template <class U>
__device__ __noinline__ void bus_unit_op(const Ctx &ctx, const BusCmd &c, const StatePtr sp, int w0, int *scr, int *out,
                                         unsigned seed, bool seeded) {
    U u;
    if (c.op == BUS_U_INIT) {
        u.load(sp, w0);
        u.init(ctx, c.value, (unsigned)c.start);
        u.store(sp, w0);
        return;
    }
    u.load(sp, w0);
    if (c.op == BUS_U_WRITE) {
        u.write(ctx, c.reg, c.value, c.start, c.dur);
        u.store(sp, w0);
        return;
    }
    const bool add = c.add & 1, wire = (c.add & 2) != 0;
    u.prepare(ctx, c.frames);
    if (seeded) u.seed(seed);
    for (int i = 0; i < c.frames; ++i) {
        int *s = scr + (size_t)(c.frame + i) * 2;
        const int in0 = s[0], in1 = s[1];
        int s0 = in0, s1 = in1, o0 = 0, o1 = 0;
        u.sample(ctx, s0, s1, o0, o1);
        if (wire) {
            int *o = out + (size_t)(c.frame + i) * 2;
            atomicAdd(o, s0);
            if (c.nout == 2) atomicAdd(o + 1, s1);
        } else if (add) {
            s[0] = wadd(in0, s0);
            if (c.nout == 2) s[1] = wadd(in1, s1);
        } else {
            s[0] = s0;
            if (c.nout == 2) s[1] = s1;
        }
    }
    u.finish();
    u.store(sp, w0);
}

__device__ __noinline__ void bus_unit_dispatch(const Ctx &ctx, const BusCmd &c, const StatePtr sp, int w0, int *scr,
                                               int *out, unsigned seed, bool seeded) {
    switch (c.kind) {
    case 1: bus_unit_op<WtOsc<false, false>>(ctx, c, sp, w0, scr, out, seed, seeded); break;
    case 2:     // panmix normally takes the BUS_PM path; kept for completeness
        if (c.nin == 1 && c.nout == 1) bus_unit_op<PanMix<1, 1, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded);
        else if (c.nin == 1) bus_unit_op<PanMix<1, 2, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded);
        else if (c.nout == 1) bus_unit_op<PanMix<2, 1, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded);
        else bus_unit_op<PanMix<2, 2, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded);
        break;
    case 3:
        if (c.nin == 1) bus_unit_op<Filter12<1, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded);
        else bus_unit_op<Filter12<2, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded);
        break;
    case 4:
        if (c.nin == 1) bus_unit_op<WaveShaper<1, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded);
        else bus_unit_op<WaveShaper<2, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded);
        break;
    case 6:
        if (c.nin == 1) bus_unit_op<Limiter<1, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded);
        else bus_unit_op<Limiter<2, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded);
        break;
    case 7:
        if (c.nin == 1) bus_unit_op<DcBlock<1, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded);
        else bus_unit_op<DcBlock<2, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded);
        break;
    case 8:
        if (c.nout == 1) bus_unit_op<Dc<1, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded);
        else bus_unit_op<Dc<2, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded);
        break;
    case 16: bus_unit_op<Fm<1, 0, 0, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded); break;
    case 17: bus_unit_op<Fm<2, 1, 0, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded); break;
    case 18: bus_unit_op<Fm<3, 2, 0, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded); break;
    case 19: bus_unit_op<Fm<4, 2, 0, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded); break;
    case 20: bus_unit_op<Fm<3, 2, 1, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded); break;
    case 21: bus_unit_op<Fm<4, 2, 1, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded); break;
    case 22: bus_unit_op<Fm<2, 1, 2, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded); break;
    case 23: bus_unit_op<Fm<4, 2, 2, false, false>>(ctx, c, sp, w0, scr, out, seed, seeded); break;
    default: break;
    }
}

// ---------------------------------------------------------------------------
// render_generic: any voice structure, one thread per voice.
//
// The fused kernels (render_bank<Chain>, render_split) exist for the structures that carry the
// load; every OTHER struct the compiler accepts (src/compiler.c:2991-3188: any list of units with
// 0-2 scratch channels between them) runs here, the way the reference itself runs a voice: unit by
// unit over each Process() segment (src/core.c:1875-1876), the scratch channels of the voice
// (st->scratch[nest], core.c:364-395) in a private [64][2] row, wire-out units adding into the
// voice's bus with integer atomics. The unit code is the same templates as everywhere else,
// dispatched at run time (bus_unit_dispatch). Same state layout, event records and segment rules
// as render_bank, in bank mode and in drop-in (explicit EV_PROC) mode.
// ---------------------------------------------------------------------------
constexpr int kMaxChain = 12;
struct GenericChain {
    int n;
    int kind[kMaxChain], nin[kMaxChain], nout[kMaxChain];
    int add[kMaxChain];         // bit 0 A2_PROCADD, bit 1 wire-out
    int word[kMaxChain];        // first state word of the unit (after the voice's flags word)
};

__global__ void __launch_bounds__(kThreads) render_generic(const RenderParams P, const GenericChain G, int *scratch) {
    const int idx = blockIdx.x * kThreads + threadIdx.x;
    if (idx >= P.nvoices) return;
    const bool expl = P.explicit_ != 0;
    const int v = P.runs ? P.runs[idx].slot : idx;
    Ctx c;
    c.waves = P.waves; c.pool = P.pool; c.cpool = P.cpool; c.ptab = P.ptab; c.fmsine = P.fmsine; c.f12tab = P.f12tab;
    c.samplerate = P.samplerate;
    const StatePtr sp{P.state + v, P.stride};
    int *scr = scratch + (size_t)v * kMaxFrag * 2;
    int alive = sp.ld(0) & 1;
    int mybus = expl ? -1 : P.bus_of[v];
    unsigned evp = 0, eve = 0;
    if (P.runs) { evp = P.runs[idx].ev_begin; eve = evp + P.runs[idx].ev_count; }
    else if (P.ev_off) { evp = P.ev_off[v]; eve = P.ev_off[v + 1]; }
    int next_ev = evp < eve ? (int)(P.ev[evp].x >> 8) : 0x7fffffff;
    unsigned seeds[kMaxChain];
    unsigned seeded = 0;
    const int W = P.W;
    BusCmd cmd;
    auto unit_cmd = [&](int op, int u) {
        cmd.op = op; cmd.kind = G.kind[u]; cmd.nin = G.nin[u]; cmd.nout = G.nout[u]; cmd.add = G.add[u];
    };
    auto apply = [&](const uint4 e) -> int {        // returns the frame count of an EV_PROC, else 0
        const int kind = e.y & 0xff, unit = (e.y >> 8) & 0xff;
        switch (kind) {
        case EV_WRITE:
        case EV_INIT:
            if (unit < G.n) {
                unit_cmd(kind == EV_INIT ? BUS_U_INIT : BUS_U_WRITE, unit);
                cmd.reg = (e.y >> 16) & 0xff; cmd.value = (int)e.z; cmd.start = (int)(e.x & 0xff); cmd.dur = (int)e.w;
                bus_unit_dispatch(c, cmd, sp, 1 + G.word[unit], nullptr, nullptr, 0, false);
            }
            break;
        case EV_START: alive = 1; break;
        case EV_STOP: alive = 0; break;
        case EV_SEED: if (unit < G.n) { seeds[unit] = e.z; seeded |= 1u << unit; } break;
        case EV_PROC: mybus = (int)e.z; return (e.y >> 8) & 0xff;
        default: break;
        }
        return 0;
    };
    for (int f0 = 0; f0 < W;) {
        const int fe = frag_end(f0, P.buffer, W);
        int f = f0;
        while (f < fe) {
            int proc_n = 0;
            while (next_ev <= f && !proc_n) {
                proc_n = apply(P.ev[evp]);
                ++evp;
                next_ev = evp < eve ? (int)(P.ev[evp].x >> 8) : 0x7fffffff;
            }
            int nxt;
            bool in_seg;
            if (expl) {
                nxt = proc_n ? min(fe, f + proc_n) : min(fe, next_ev);
                in_seg = proc_n != 0;
            } else {
                nxt = min(fe, next_ev);
                for (int k = 0; k < P.nsplits; ++k)
                    if (P.splits[k] > f) nxt = min(nxt, P.splits[k]);
                in_seg = alive != 0;
            }
            if (in_seg && mybus >= 0) {
                int *out = P.acc + ((size_t)mybus * W + f0) * 2;
                for (int u = 0; u < G.n; ++u) {         // core.c:1875-1876
                    unit_cmd(BUS_U_RUN, u);
                    cmd.frame = f - f0; cmd.frames = nxt - f;
                    bus_unit_dispatch(c, cmd, sp, 1 + G.word[u], scr, out, seeds[u], (seeded >> u) & 1);
                }
            }
            seeded = 0;
            f = nxt;
        }
        f0 = fe;
    }
    if (expl)       // writes after the voice's last segment of the fragment (see render_bank)
        while (evp < eve) apply(P.ev[evp++]);
    sp.st(0, alive);
}

// fbdelay (units/fbdelay.c). State words: 0 fbdelay, 1 ldelay, 2 rdelay (frames,
// converted on the host, fbdelay.c:229-245), 3 drygain, 4 fbgain, 5 lgain,
// 6 rgain (16:16), 7 bufpos, 8/9 device pointer of the two delay lines
// [2][kFbdSize] (zeroed by the host at allocation, fbdelay.c:187-188).
A2CU_DEV void fbd_frame(const BusCmd &c, const int *st, int *b0, int *b1, int *acc, int i, bool wire, bool add) {
    const unsigned mask = kFbdSize - 1;
    const unsigned pos = (unsigned)st[7] + (unsigned)i;
    int *s = acc + ((size_t)c.in_bus * kMaxFrag + c.frame + i) * 2;
    const int i0 = s[0];
    const int i1 = c.nin == 2 ? s[1] : i0;
    // fbdelay.c:86-101 (feedback taps are cross-fed: "reverse stereo")
    int o0 = mulshr(b1[(pos - (unsigned)st[0]) & mask], st[4], 16);
    int o1 = mulshr(b0[(pos - (unsigned)st[0]) & mask], st[4], 16);
    b0[pos & mask] = wadd(i0, o0);
    b1[pos & mask] = wadd(i1, o1);
    o0 = wadd(o0, mulshr(b0[(pos - (unsigned)st[1]) & mask], st[5], 16));
    o1 = wadd(o1, mulshr(b1[(pos - (unsigned)st[2]) & mask], st[6], 16));
    o0 = wadd(o0, mulshr(i0, st[3], 16));
    o1 = wadd(o1, mulshr(i1, st[3], 16));
    if (c.nout == 1) { o0 = wadd(o0, o1) >> 1; o1 = 0; }        // fbdelay.c:110, 119
    if (wire) {
        int *o = acc + ((size_t)c.out_bus * kMaxFrag + c.frame + i) * 2;
        atomicAdd(o, o0);
        if (c.nout == 2) atomicAdd(o + 1, o1);
    } else if (add) {
        s[0] = wadd(s[0], o0);
        if (c.nout == 2) s[1] = wadd(s[1], o1);
    } else {
        s[0] = o0;
        if (c.nout == 2) s[1] = o1;
    }
}

A2CU_DEV void fbd_op(const BusCmd &c, int *st, int *acc, int tid) {
    if (c.op == BUS_U_INIT) {
        if (tid == 0) {
            for (int i = 0; i < 8; ++i) st[i] = 0;
            st[8] = c.value; st[9] = c.dur;         // delay-line pointer, low / high word
        }
        return;
    }
    if (c.op == BUS_U_WRITE) {
        if (tid == 0 && c.reg >= 0 && c.reg < 7) st[c.reg] = c.value;
        return;
    }
    int *b0 = (int *)(((unsigned long long)(unsigned)st[9] << 32) | (unsigned)st[8]);
    int *b1 = b0 + kFbdSize;
    const bool add = c.add & 1, wire = (c.add & 2) != 0;
    const unsigned mask = kFbdSize - 1;
    bool par = true;        // no tap of this call reads a slot this call writes
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned d = (unsigned)st[k] & mask;
        par = par && d >= (unsigned)c.frames && d <= (unsigned)(kFbdSize - c.frames);
    }
    if (par) {
        if (tid < c.frames) fbd_frame(c, st, b0, b1, acc, tid, wire, add);
    } else if (tid == 0) {
        for (int i = 0; i < c.frames; ++i) fbd_frame(c, st, b0, b1, acc, i, wire, add);
    }
    __syncthreads();
    if (tid == 0) st[7] = (int)((unsigned)st[7] + (unsigned)c.frames);
}

__global__ void __launch_bounds__(kMaxFrag) bus_level(const BusVmParams P) {
    __shared__ MixSeg sg;
    const int tid = threadIdx.x;
    int *acc = P.acc;
    const BusRun run = P.runs[blockIdx.x];
    unsigned seed = 0;
    bool seeded = false;
    for (unsigned ci = run.begin; ci < run.begin + run.count; ++ci) {
        const BusCmd c = P.cmds[ci];
        if (c.op >= BUS_U_INIT && c.op <= BUS_U_RUN) {
            if (c.op == BUS_U_SEED) { seed = (unsigned)c.value; seeded = true; continue; }
            int *ust = P.ustate + (size_t)c.pm * kUnitWords;
            if (c.kind == kFbdKind) fbd_op(c, ust, acc, tid);
            else if (tid == 0)
                bus_unit_dispatch(P.ctx, c, StatePtr{ust, 1}, 0, acc + (size_t)c.in_bus * kMaxFrag * 2,
                                  acc + (size_t)c.out_bus * kMaxFrag * 2, seed, seeded);
            if (c.op == BUS_U_RUN) seeded = false;
            __syncthreads();
            continue;
        }
        if (c.op == BUS_ADD) {
            if (tid < c.frames) {
                const int *in = acc + ((size_t)c.in_bus * kMaxFrag + c.frame + tid) * 2;
                int *out = acc + ((size_t)c.out_bus * kMaxFrag + c.frame + tid) * 2;
                atomicAdd(out, in[0]);
                atomicAdd(out + 1, in[1]);
            }
            __syncthreads();
            continue;
        }
        int *st = P.pmstate + (size_t)c.pm * 8;
        if (c.op == BUS_PM_WRITE) {
            if (tid == 0) {
                Ramp vol, pan;
                pm_load(st, vol, pan);
                ramp_set(c.reg == 0 ? vol : pan, c.value, c.start, c.dur);
                pm_store(st, vol, pan);
            }
            __syncthreads();
            continue;
        }
        if (tid == 0) {
            Ramp vol, pan;
            pm_load(st, vol, pan);
            const bool one = c.nin == 1 && c.nout == 1;     // panmix.c:49-64
            sg.clamp = !one && (pan.target > 0xffffff || pan.target < -0xffffff ||
                                pan.value > 0xffffff || pan.value < -0xffffff);
            ramp_prepare(vol, c.frames);
            if (!one) ramp_prepare(pan, c.frames);
            sg.vol = vol.value; sg.dvol = vol.delta;
            sg.pan = pan.value; sg.dpan = one ? 0 : pan.delta;
            ramp_run(vol, c.frames);
            if (!one) ramp_run(pan, c.frames);
            pm_store(st, vol, pan);
        }
        __syncthreads();
        if (tid < c.frames) {
            const int f = c.frame + tid;
            const int *in = acc + ((size_t)c.in_bus * kMaxFrag + f) * 2;
            int *out = acc + ((size_t)c.out_bus * kMaxFrag + f) * 2;
            const int i0 = in[0], i1 = in[1];
            const int v = wadd(sg.vol, wmul(sg.dvol, tid));
            int r0, r1 = 0;
            if (c.nin == 1 && c.nout == 1) {
                r0 = mulshr(i0, v, 24);
            } else {
                const int pn = wadd(sg.pan, wmul(sg.dpan, tid));
                const int vp = mulshr(pn, v, 24);
                int v0 = wsub(v, vp), v1 = wadd(v, vp);
                if (sg.clamp) {
                    const int lim = (int)((unsigned)v << 1);
                    if (v0 > lim) v0 = lim;
                    if (v1 > lim) v1 = lim;
                }
                if (c.nin == 1) { r0 = mulshr(i0, v0, 24); r1 = mulshr(i0, v1, 24); }
                else if (c.nout == 1) r0 = (int)(((long long)i0 * v0 + (long long)i1 * v1) >> 25);
                else { r0 = mulshr(i0, v0, 24); r1 = mulshr(i1, v1, 24); }
            }
            if (c.out_bus != c.in_bus) {
                // another voice's bus (wire-out) or a private row: atomics are safe in both cases
                if (c.add) { atomicAdd(out, r0); if (c.nout == 2) atomicAdd(out + 1, r1); }
                else { out[0] = r0; if (c.nout == 2) out[1] = r1; }
            } else if (c.add) { out[0] = wadd(out[0], r0); if (c.nout == 2) out[1] = wadd(out[1], r1); }
            else { out[0] = r0; if (c.nout == 2) out[1] = r1; }
        }
        __syncthreads();
    }
}

}  // namespace a2cu
