// a2cu_kernels.cuh - render kernels of the voice engine.
//
//   render_bank<Chain>   one thread = one voice, frame-synchronous over a
//                        window of W frames; replays each voice's segment list
//                        (core.c:1847-1880), runs the fused unit chain per
//                        sample and reduces the warp's voices with
//                        redux.sync (__reduce_add_sync) into a per-CTA shared
//                        memory bus, flushed to the device bus with integer
//                        atomics once per fragment.  Integer add is
//                        associative, so any reduction order is bit-exact.
//   mix_groups/mix_root  group buses -> group panmix -> root bus -> root
//                        panmix -> master (core.c:1763-1776, audiality2.c:
//                        268-304, xinsert bypass xinsert.c:149-156)
#pragma once
#include "a2cu_device.cuh"

namespace a2cu {

constexpr int kThreads = 128;
constexpr int kMaxSplits = 8;

// Event record, 16 bytes. x = (frame_in_window << 8) | substart
// y = kind | unit << 8 | reg << 16 ; z = value ; w = duration (24:8)
// EV_PROC (drop-in mode): y = kind | frames << 8, z = device bus index: one
// Process() call of the voice's units, exactly as the host walked it.
// EV_SEED: z = start state of the shared noise LCG for the next segment of unit
// 'unit' (computed on the host in tree-walk order, wtosc.c:135-144)
enum EvKind { EV_WRITE = 0, EV_WAKE = 1, EV_INIT = 2, EV_START = 3, EV_STOP = 4, EV_PROC = 5, EV_SEED = 6 };

// One active voice of a drop-in block: slot and its run of event records.
struct VoiceRun { int slot; unsigned ev_begin, ev_count; };

struct RenderParams {
    int *state;               // [words][stride]
    size_t stride;
    int nvoices;
    const int *bus_of;        // per voice: 0 = root bus, 1 + g = group g
    const unsigned *ev_off;   // CSR offsets [nvoices + 1] or nullptr
    const uint4 *ev;
    int *acc;                 // [nbus][W][2]
    int W;                    // frames in this window
    int buffer;               // driver buffer size
    int nsplits;
    int splits[kMaxSplits];   // forced split frames (root wake-ups / writes)
    const WaveDesc *waves;
    const int16_t *pool;
    const int4 *cpool;
    const unsigned *ptab;
    const int16_t *fmsine;
    int samplerate;
    // drop-in ("block") mode: thread i renders runs[i]; segments are the
    // host's explicit EV_PROC records and carry their own target bus
    const VoiceRun *runs;
    int explicit_;
    // optional per-role busy-cycle counters of render_split (debug/profiling)
    unsigned long long *prof;
    // render_split: coefficient-pool range [stage_begin, stage_begin + stage_count)
    // of the bank's wave, staged into shared memory with one TMA bulk copy
    int stage_begin, stage_count;
    // render_split: fused root stage. The last CTA to finish (ticket from fuse_counter) runs the root
    // panmix over the whole window - same code as mix_root - so the step needs no second launch.
    int fuse_root;
    unsigned *fuse_counter;
    int *fuse_rstate;
    int *fuse_master;
    int fuse_channels;
    int fuse_root_stage;    // 0: copy the raw root bus out (multi-GPU cut) instead of the root panmix
};

// End of the fragment that contains frame f: fragments restart at every driver
// buffer and are at most 64 frames (core.c:1964-1973).
A2CU_DEV int frag_end(int f, int buffer, int W) {
    int pos = f % buffer;
    int e = f - pos + min(buffer, (pos / kMaxFrag + 1) * kMaxFrag);
    return min(e, W);
}

template <class CH>
__global__ void __launch_bounds__(kThreads) render_bank(const RenderParams P) {
    __shared__ int sacc[kMaxFrag][2];
    __shared__ int16_t s_sine[CH::kUsesFm ? 2049 : 1];
    __shared__ int s_home;

    const int tid = threadIdx.x;
    const int idx = blockIdx.x * kThreads + tid;
    const bool valid = idx < P.nvoices;
    const bool expl = P.explicit_ != 0;
    const int v = (valid && P.runs) ? P.runs[idx].slot : idx;

    if (CH::kUsesFm)
        for (int i = tid; i < 2049; i += kThreads) s_sine[i] = P.fmsine[i];
    if (tid < kMaxFrag) { sacc[tid][0] = 0; sacc[tid][1] = 0; }
    int mybus = (valid && !expl) ? P.bus_of[v] : -1;
    if (tid == 0) s_home = expl ? -2 : mybus;
    __syncthreads();
    const int home = s_home;

    Ctx c;
    c.waves = P.waves; c.pool = P.pool; c.cpool = P.cpool; c.ptab = P.ptab; c.fmsine = s_sine;
    c.samplerate = P.samplerate;

    CH ch;
    StatePtr sp{P.state + (valid ? v : 0), P.stride};
    int alive = 0;
    if (valid) { alive = sp.ld(0) & 1; ch.load(sp, 1); }

    unsigned evp = 0, eve = 0;
    if (valid && P.runs) { evp = P.runs[idx].ev_begin; eve = evp + P.runs[idx].ev_count; }
    else if (valid && P.ev_off) { evp = P.ev_off[v]; eve = P.ev_off[v + 1]; }
    int next_ev = evp < eve ? (int)(P.ev[evp].x >> 8) : 0x7fffffff;

    int seg_end = 0;
    bool in_seg = false;
    const int W = P.W;

    for (int f0 = 0; f0 < W;) {
        const int fe = frag_end(f0, P.buffer, W);
        // Bus mix-down of one frame (PROCADD "+=", wtosc.c:229, panmix.c:104-105):
        // integer adds in any order. Bank mode: voices at the CTA's home bus are
        // summed with redux.sync into the shared-memory bus. Drop-in mode: each
        // segment names its own bus, so the warp sums the lanes that share the
        // first active lane's bus; stragglers add directly.
        auto accumulate = [&](int fr, bool athome, int o0, int o1) {
            if (!expl) {
                int h0 = __reduce_add_sync(0xffffffffu, athome ? o0 : 0);
                int h1 = __reduce_add_sync(0xffffffffu, athome ? o1 : 0);
                if ((tid & 31) == 0) {
                    atomicAdd(&sacc[fr - f0][0], h0);
                    atomicAdd(&sacc[fr - f0][1], h1);
                }
                if (valid && !athome && in_seg) {
                    int *a = P.acc + ((size_t)mybus * W + fr) * 2;
                    atomicAdd(a, o0);
                    atomicAdd(a + 1, o1);
                }
                return;
            }
            const unsigned act = __ballot_sync(0xffffffffu, in_seg);
            if (!act) return;
            const int leader = __ffs(act) - 1;
            const int lbus = __shfl_sync(0xffffffffu, mybus, leader);
            const bool same = in_seg && mybus == lbus;
            int h0 = __reduce_add_sync(0xffffffffu, same ? o0 : 0);
            int h1 = __reduce_add_sync(0xffffffffu, same ? o1 : 0);
            if ((tid & 31) == leader) {
                int *a = P.acc + ((size_t)lbus * W + fr) * 2;
                if (h0) atomicAdd(a, h0);
                if (h1) atomicAdd(a + 1, h1);
            }
            if (in_seg && !same) {
                int *a = P.acc + ((size_t)mybus * W + fr) * 2;
                atomicAdd(a, o0);
                atomicAdd(a + 1, o1);
            }
        };
        int f = f0;
        // Segment boundary work for the first frame of the fragment
        auto boundary = [&](int fr) {
            if (in_seg) ch.finish();
            int proc_n = 0;
            while (next_ev <= fr && !proc_n) {
                const uint4 e = P.ev[evp];
                const int kind = e.y & 0xff, unit = (e.y >> 8) & 0xff;
                const int reg = (e.y >> 16) & 0xff;
                switch (kind) {
                case EV_WRITE: ch.write(c, unit, reg, (int)e.z, (int)(e.x & 0xff), (int)e.w); break;
                case EV_INIT: ch.init_unit(c, unit, (int)e.z, e.x & 0xff); break;
                case EV_START: alive = 1; break;
                case EV_STOP: alive = 0; break;
                case EV_SEED: ch.seed_unit(unit, e.z); break;
                case EV_PROC: proc_n = (e.y >> 8) & 0xff; mybus = (int)e.z; break;
                default: break;
                }
                ++evp;
                next_ev = evp < eve ? (int)(P.ev[evp].x >> 8) : 0x7fffffff;
            }
            int nxt;
            if (expl) {
                // the host decided the segment: [fr, fr + proc_n), or idle
                nxt = proc_n ? min(fe, fr + proc_n) : min(fe, next_ev);
                in_seg = proc_n != 0;
            } else {
                nxt = min(fe, next_ev);
                for (int k = 0; k < P.nsplits; ++k)
                    if (P.splits[k] > fr) nxt = min(nxt, P.splits[k]);
                in_seg = alive != 0;
            }
            seg_end = nxt;
            if (in_seg) ch.prepare(c, nxt - fr);
        };
        if (valid && f == seg_end) boundary(f);
        const bool athome0 = mybus == home;
        // Fast path: every voice of the warp runs one plain segment across the
        // whole fragment -> straight-line sample loop, unrolled for ILP.
        const bool ok = !valid || (seg_end == fe && (!in_seg || ch.plain()));
        if (__all_sync(0xffffffffu, ok)) {
#pragma unroll 4
            for (; f < fe; ++f) {
                int s0 = 0, s1 = 0, o0 = 0, o1 = 0;
                if (in_seg) ch.sample_fast(c, s0, s1, o0, o1);
                accumulate(f, athome0, o0, o1);
            }
        }
        for (; f < fe; ++f) {
            if (valid && f == seg_end && f != f0) boundary(f);
            int s0 = 0, s1 = 0, o0 = 0, o1 = 0;
            if (in_seg) ch.sample(c, s0, s1, o0, o1);
            accumulate(f, mybus == home, o0, o1);
        }
        __syncthreads();
        if (tid < (fe - f0) * 2 && home >= 0) {
            int val = sacc[tid >> 1][tid & 1];
            if (val) atomicAdd(P.acc + ((size_t)home * W + f0) * 2 + tid, val);
            sacc[tid >> 1][tid & 1] = 0;
        }
        __syncthreads();
        f0 = fe;
    }
    if (valid) {
        if (in_seg) ch.finish();
        sp.st(0, alive);
        ch.store(sp, 1);
    }
}

// ---------------------------------------------------------------------------
// Bus stage
// ---------------------------------------------------------------------------
struct MixEvent {           // 16 bytes
    unsigned time;          // (frame_in_window << 8) | substart
    int target;             // group index, or -1 for the root
    int reg_dur_hi;         // reg (int8, -1 = wake) in low 8 bits
    int value;
    unsigned dur;
    int pad[3];
};

struct MixParams {
    int *acc;               // [nbus][W][2]; bus 0 = root
    int W, buffer, ngroups, channels;
    int nsplits;
    int splits[kMaxSplits];
    int *gstate;            // [ngroups][8] panmix rampers (vol, pan)
    int *rstate;            // [8]
    const MixEvent *ev;     // sorted by time
    int nev;
    int *master;            // [W][channels], or raw root bus [W][2] if !root_stage (may be mapped host memory)
    int root_stage;
    int clear;              // consumers zero the bus rows they read, so the next window needs no memset
};

A2CU_DEV void pm_load(const int *s, Ramp &vol, Ramp &pan) {
    vol.value = s[0]; vol.target = s[1]; vol.delta = s[2]; vol.timer = s[3];
    pan.value = s[4]; pan.target = s[5]; pan.delta = s[6]; pan.timer = s[7];
}
A2CU_DEV void pm_store(int *s, const Ramp &vol, const Ramp &pan) {
    s[0] = vol.value; s[1] = vol.target; s[2] = vol.delta; s[3] = vol.timer;
    s[4] = pan.value; s[5] = pan.target; s[6] = pan.delta; s[7] = pan.timer;
}

// One bus-level panmix (panmix.c:137-229) over the window, one CTA per bus.
// Thread 0 replays the control-rate part (events, a2_PrepareRamper per
// segment) into per-segment records; the ramps are linear inside a segment,
// so every frame is then evaluated independently by the whole CTA.
struct MixSeg {
    int f0, f1;         // [f0, f1)
    int vol, dvol, pan, dpan;
    int clamp;
};
constexpr int kMixSegs = 128;

template <class Emit>
A2CU_DEV void pm_bus(const MixParams &P, int target, int *state, const int *in, bool mono, Emit emit) {
    __shared__ MixSeg seg[kMixSegs];
    __shared__ int s_nseg, s_next, s_evp;
    __shared__ int s_state[8];
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_next = 0; s_evp = 0;
        for (int i = 0; i < 8; ++i) s_state[i] = state[i];
    }
    __syncthreads();
    while (true) {
        const int c0 = s_next;
        if (c0 >= P.W) break;
        if (tid == 0) {
            Ramp vol, pan;
            pm_load(s_state, vol, pan);
            int evp = s_evp;
            auto next_time = [&]() {
                while (evp < P.nev && P.ev[evp].target != target) ++evp;
                return evp < P.nev ? (int)(P.ev[evp].time >> 8) : 0x7fffffff;
            };
            int next_ev = next_time();
            int f = c0, n = 0;
            while (f < P.W && n < kMixSegs) {
                while (next_ev <= f) {
                    const MixEvent &e = P.ev[evp];
                    int reg = (int)(signed char)(e.reg_dur_hi & 0xff);
                    if (reg >= 0) ramp_set(reg == 0 ? vol : pan, e.value, (int)(e.time & 0xff), (int)e.dur);
                    ++evp;
                    next_ev = next_time();
                }
                int nxt = min(frag_end(f, P.buffer, P.W), next_ev);
                // a group's segments are additionally cut by its parent's (the root's)
                for (int k = 0; k < P.nsplits; ++k)
                    if (P.splits[k] > f) nxt = min(nxt, P.splits[k]);
                MixSeg sg;
                sg.clamp = pan.target > 0xffffff || pan.target < -0xffffff ||
                           pan.value > 0xffffff || pan.value < -0xffffff;
                ramp_prepare(vol, nxt - f);
                ramp_prepare(pan, nxt - f);
                sg.f0 = f; sg.f1 = nxt;
                sg.vol = vol.value; sg.dvol = vol.delta; sg.pan = pan.value; sg.dpan = pan.delta;
                seg[n++] = sg;
                ramp_run(vol, nxt - f);
                ramp_run(pan, nxt - f);
                f = nxt;
            }
            pm_store(s_state, vol, pan);
            s_nseg = n; s_next = f; s_evp = evp;
        }
        __syncthreads();
        const int c1 = s_next, nseg = s_nseg;
        for (int f = c0 + tid; f < c1; f += blockDim.x) {
            int lo = 0, hi = nseg - 1;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (seg[mid].f1 <= f) lo = mid + 1; else hi = mid;
            }
            const MixSeg sg = seg[lo];
            int k = f - sg.f0;
            int v = wadd(sg.vol, wmul(sg.dvol, k));
            int pn = wadd(sg.pan, wmul(sg.dpan, k));
            int vp = mulshr(pn, v, 24);
            int v0 = wsub(v, vp), v1 = wadd(v, vp);
            if (sg.clamp) {
                int lim = (int)((unsigned)v << 1);
                if (v0 > lim) v0 = lim;
                if (v1 > lim) v1 = lim;
            }
            int i0 = __ldcg(in + f * 2), i1 = __ldcg(in + f * 2 + 1);
            if (mono)
                emit(f, (int)(((long long)i0 * v0 + (long long)i1 * v1) >> 25), 0);
            else
                emit(f, mulshr(i0, v0, 24), mulshr(i1, v1, 24));
        }
        __syncthreads();
    }
    if (tid == 0)
        for (int i = 0; i < 8; ++i) state[i] = s_state[i];
}

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
// Root stage over the window for a group of threads (gtid of gsize); 'cta0' tells whether this CTA
// runs the general (segment replay) path when the rampers are not at rest.
A2CU_DEV void root_stage(const MixParams &P, int gtid, int gsize, bool cta0) {
    int *root = P.acc;
    const bool clear = P.clear != 0;
    if (!P.root_stage) {
        for (int i = gtid; i < P.W * 2; i += gsize) { P.master[i] = root[i]; if (clear) root[i] = 0; }
        return;
    }
    const bool mono = P.channels == 1;
    // Steady state (no root events or splits in the window, both rampers at
    // rest): a2_PrepareRamper leaves value = target, delta = 0 in every segment
    // (a2_dsp.h:130-134), so every frame uses the same two gains and the whole
    // grid evaluates frames independently. Otherwise CTA 0 replays the segments.
    const bool steady = P.nev == 0 && P.nsplits == 0 && P.rstate[3] == 0 && P.rstate[7] == 0 &&
                        P.rstate[0] == P.rstate[1] && P.rstate[4] == P.rstate[5];
    if (steady) {
        const int v = P.rstate[1], pn = P.rstate[5];            // targets
        const int vp = mulshr(pn, v, 24);
        int v0 = wsub(v, vp), v1 = wadd(v, vp);
        if (pn > 0xffffff || pn < -0xffffff) {
            const int lim = (int)((unsigned)v << 1);            // panmix.c:117-135 clamp variant
            if (v0 > lim) v0 = lim;
            if (v1 > lim) v1 = lim;
        }
        for (int f = gtid; f < P.W; f += gsize) {
            const int i0 = __ldcg(root + f * 2), i1 = __ldcg(root + f * 2 + 1);
            if (clear) { root[f * 2] = 0; root[f * 2 + 1] = 0; }
            if (mono) P.master[f] = (int)(((long long)i0 * v0 + (long long)i1 * v1) >> 25);
            else { P.master[f * 2] = mulshr(i0, v0, 24); P.master[f * 2 + 1] = mulshr(i1, v1, 24); }
        }
        if (gtid == 0 && P.W > 0) { P.rstate[2] = 0; P.rstate[6] = 0; }     // deltas as PrepareRamper leaves them
        return;
    }
    if (!cta0) return;
    pm_bus(P, -1, P.rstate, root, mono, [&](int f, int r0, int r1) {
        if (mono) P.master[f] = r0;
        else { P.master[f * 2] = r0; P.master[f * 2 + 1] = r1; }
        if (clear) { root[f * 2] = 0; root[f * 2 + 1] = 0; }
    });
}

// root: { inline 0 *|2; panmix * *|2 1; xinsert * > } into the cleared master;
// with root_stage == 0 the raw root bus is copied out (multi-GPU cut).
__global__ void __launch_bounds__(256) mix_root(const MixParams P) {
    root_stage(P, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, blockIdx.x == 0);
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

template <class U>
__device__ __noinline__ void bus_unit_op(const Ctx &ctx, const BusCmd &c, int *st, int *acc, unsigned seed, bool seeded) {
    U u;
    const StatePtr sp{st, 1};
    if (c.op == BUS_U_INIT) {
        u.load(sp, 0);
        u.init(ctx, c.value, (unsigned)c.start);
        u.store(sp, 0);
        return;
    }
    u.load(sp, 0);
    if (c.op == BUS_U_WRITE) {
        u.write(ctx, c.reg, c.value, c.start, c.dur);
        u.store(sp, 0);
        return;
    }
    const bool add = c.add & 1, wire = (c.add & 2) != 0;
    u.prepare(ctx, c.frames);
    if (seeded) u.seed(seed);
    for (int i = 0; i < c.frames; ++i) {
        int *s = acc + ((size_t)c.in_bus * kMaxFrag + c.frame + i) * 2;
        const int in0 = s[0], in1 = s[1];
        int s0 = in0, s1 = in1, o0 = 0, o1 = 0;
        u.sample(ctx, s0, s1, o0, o1);
        if (wire) {
            int *o = acc + ((size_t)c.out_bus * kMaxFrag + c.frame + i) * 2;
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
    u.store(sp, 0);
}

__device__ __noinline__ void bus_unit_dispatch(const Ctx &ctx, const BusCmd &c, int *st, int *acc, unsigned seed,
                                               bool seeded) {
    switch (c.kind) {
    case 1: bus_unit_op<WtOsc<false, false>>(ctx, c, st, acc, seed, seeded); break;
    case 2:     // panmix normally takes the BUS_PM path; kept for completeness
        if (c.nin == 1 && c.nout == 1) bus_unit_op<PanMix<1, 1, false, false>>(ctx, c, st, acc, seed, seeded);
        else if (c.nin == 1) bus_unit_op<PanMix<1, 2, false, false>>(ctx, c, st, acc, seed, seeded);
        else if (c.nout == 1) bus_unit_op<PanMix<2, 1, false, false>>(ctx, c, st, acc, seed, seeded);
        else bus_unit_op<PanMix<2, 2, false, false>>(ctx, c, st, acc, seed, seeded);
        break;
    case 3:
        if (c.nin == 1) bus_unit_op<Filter12<1, false, false>>(ctx, c, st, acc, seed, seeded);
        else bus_unit_op<Filter12<2, false, false>>(ctx, c, st, acc, seed, seeded);
        break;
    case 4:
        if (c.nin == 1) bus_unit_op<WaveShaper<1, false, false>>(ctx, c, st, acc, seed, seeded);
        else bus_unit_op<WaveShaper<2, false, false>>(ctx, c, st, acc, seed, seeded);
        break;
    case 16: bus_unit_op<Fm<1, 0, 0, false, false>>(ctx, c, st, acc, seed, seeded); break;
    case 17: bus_unit_op<Fm<2, 1, 0, false, false>>(ctx, c, st, acc, seed, seeded); break;
    case 18: bus_unit_op<Fm<3, 2, 0, false, false>>(ctx, c, st, acc, seed, seeded); break;
    case 19: bus_unit_op<Fm<4, 2, 0, false, false>>(ctx, c, st, acc, seed, seeded); break;
    case 20: bus_unit_op<Fm<3, 2, 1, false, false>>(ctx, c, st, acc, seed, seeded); break;
    case 21: bus_unit_op<Fm<4, 2, 1, false, false>>(ctx, c, st, acc, seed, seeded); break;
    case 22: bus_unit_op<Fm<2, 1, 2, false, false>>(ctx, c, st, acc, seed, seeded); break;
    case 23: bus_unit_op<Fm<4, 2, 2, false, false>>(ctx, c, st, acc, seed, seeded); break;
    default: break;
    }
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
            else if (tid == 0) bus_unit_dispatch(P.ctx, c, ust, acc, seed, seeded);
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
