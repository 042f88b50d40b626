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
constexpr int kSplitSegs = 2;      // render_split: segments per voice and fragment (host eligibility check)
constexpr int kSplitMaxWin = 1024;  // render_split: frames per launch (per-CTA bus accumulator in shared memory)

// Event record, 16 bytes. x = (frame_in_window << 8) | substart
// y = kind | unit << 8 | reg << 16 ; z = value ; w = duration (24:8)
// EV_PROC (drop-in mode): y = kind | frames << 8, z = device bus index: one
// Process() call of the voice's units, exactly as the host walked it.
// EV_SEED: z = start state of the shared noise LCG for the next segment of unit
// 'unit' (computed on the host in tree-walk order, wtosc.c:135-144)
enum EvKind { EV_WRITE = 0, EV_WAKE = 1, EV_INIT = 2, EV_START = 3, EV_STOP = 4, EV_PROC = 5, EV_SEED = 6 };

// One active voice of a drop-in block: slot and its run of event records.
struct VoiceRun { int slot; unsigned ev_begin, ev_count; };

// ---------------------------------------------------------------------------
// Multi-GPU exchange of the root bus over NVLink peer memory.
//
// Voices shard across GPUs; the only coupling is the integer sum of the stereo root
// scratch bus BEFORE the truncating root panmix (audiality2.c:271-280). Every rank owns
// a "symmetric" buffer  flags[2][world] | data[2][world][max_frames][2]  that all peers
// have mapped (CUDA IPC / peer access). At the end of a window the CTA that holds the
// finished root bus
//   1. stores its raw bus into row [epoch & 1][rank] of EVERY rank's buffer (plain
//      stores through the peer mapping: NVLink writes), fences system-wide and
//      releases flag [epoch & 1][rank] = epoch on every rank,
//   2. waits until all `world` flags of its OWN buffer show this epoch,
//   3. sums the rows (integer add: any order is bit-exact) back into its root bus,
// and goes on with the ordinary root stage. No launch, no NCCL call and no host
// round trip is involved; two buffer halves alternate so that a rank that is one
// window ahead never overwrites rows a slower rank still reads (a rank cannot get two
// windows ahead: finishing window k+1 needs every peer's push of k+1, which peers
// issue only after they finished reading window k).
// The wait is bounded (timeout -> status word), so a missing peer cannot hang the GPU.
// ---------------------------------------------------------------------------
constexpr int kMaxPeers = 8;
struct XchgParams {
    int world, rank;                    // world <= 1: no exchange
    unsigned epoch;                     // sequence number of this window (> 0, same on all ranks)
    int max_frames;                     // row pitch of the data area, frames
    int *data[kMaxPeers];               // every rank's buffer: data area
    unsigned *flags[kMaxPeers];         // every rank's buffer: flag area
    unsigned *status;                   // mapped host word: set to the epoch that timed out
    unsigned long long timeout_cycles;
    // Lagged mode (pipelined a2cu_submit / a2cu_collect): this launch only PUBLISHES its root bus;
    // the sum over the ranks and the root stage of the PREVIOUS window run here instead, when its
    // rows have long arrived - the NVLink round trips (stores, system fence, flag, poll) leave the
    // critical path, at the price of one window of output latency.
    int lag;
    int prev_valid;                     // a previous window is waiting for its root stage
    unsigned prev_epoch;
    int prev_W, prev_buffer, prev_nsplits;
    int prev_splits[kMaxSplits];
    int *prev_master;                   // where the previous window's master block goes
    int *sum;                           // [max_frames][2] scratch: summed root bus of the previous window
};

A2CU_DEV void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
A2CU_DEV unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All threads of ONE CTA (nthreads of them) call these with the finished local root bus.
// Both are latency-bound single-CTA passes over a few KB, so every thread first issues ALL its
// loads (four per round, independent), then stores; and the release is done ONCE, by the
// flag-writing threads: st.release.sys after the CTA barrier is cumulative over the data stores the
// barrier ordered before it (a membar.sys in each of 512 threads cost ~5 us, profiles/xchg_tail.py).
// Publish: raw bus -> row [epoch & 1][rank] of every rank's buffer, then the flag.
A2CU_DEV void xchg_publish(const XchgParams &X, int *root, int W, int tid, int nthreads, bool zero) {
    const int par = (int)(X.epoch & 1u);
    const size_t pitch = (size_t)X.max_frames * 2;
    const size_t myrow = ((size_t)par * X.world + X.rank) * pitch;
    const int n = W * 2;
    for (int i0 = tid; i0 < n; i0 += 4 * nthreads) {
        int v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = i0 + j * nthreads < n ? __ldcg(root + i0 + j * nthreads) : 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = i0 + j * nthreads;
            if (i < n) {
                for (int r = 0; r < X.world; ++r) X.data[r][myrow + i] = v[j];
                if (zero) root[i] = 0;      // lagged mode: nobody else clears the local bus rows
            }
        }
    }
    __syncthreads();
    if (tid < X.world) st_release_sys(X.flags[tid] + par * X.world + X.rank, X.epoch);
}
// Collect: wait until every rank's flag shows `epoch`, sum the rows into dst[W][2].
A2CU_DEV void xchg_collect(const XchgParams &X, unsigned epoch, int *dst, int W, int tid, int nthreads) {
    const int par = (int)(epoch & 1u);
    const size_t pitch = (size_t)X.max_frames * 2;
    if (tid < X.world) {
        const unsigned *f = X.flags[X.rank] + par * X.world + tid;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
            if ((unsigned long long)(clock64() - t0) > X.timeout_cycles) {
                *X.status = epoch;          // reported by a2cu_collect / a2cu_sync
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    const int *mine = X.data[X.rank] + (size_t)par * X.world * pitch;
    const int n = W * 2;
    for (int i0 = tid; i0 < n; i0 += 4 * nthreads) {
        int sacc[4] = {0, 0, 0, 0};
        for (int r = 0; r < X.world; ++r) {
            int v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = i0 + j * nthreads < n ? __ldcg(mine + (size_t)r * pitch + i0 + j * nthreads) : 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) sacc[j] = wadd(sacc[j], v[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i0 + j * nthreads < n) dst[i0 + j * nthreads] = sacc[j];
    }
    __syncthreads();
}
A2CU_DEV void xchg_root_bus(const XchgParams &X, int *root, int W, int tid, int nthreads) {
    xchg_publish(X, root, W, tid, nthreads, false);
    xchg_collect(X, X.epoch, root, W, tid, nthreads);
}

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
    const int *fmsine;        // packed FM sine, Ctx::fmsine
    const int *f12tab;
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
    int fuse_out_fmt;       // MixParams::out_fmt of the fused root stage
    XchgParams xchg;        // world > 1: sum the root bus over all ranks before the root stage
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
    __shared__ int s_sine[CH::kUsesFm ? 2048 : 1];
    __shared__ int s_home;

    const int tid = threadIdx.x;
    const int idx = blockIdx.x * kThreads + tid;
    const bool valid = idx < P.nvoices;
    const bool expl = P.explicit_ != 0;
    const int v = (valid && P.runs) ? P.runs[idx].slot : idx;

    if (CH::kUsesFm)
        for (int i = tid; i < 2048; i += kThreads) s_sine[i] = P.fmsine[i];
    if (tid < kMaxFrag) { sacc[tid][0] = 0; sacc[tid][1] = 0; }
    int mybus = (valid && !expl) ? P.bus_of[v] : -1;
    if (tid == 0) s_home = expl ? -2 : mybus;
    __syncthreads();
    const int home = s_home;

    Ctx c;
    c.waves = P.waves; c.pool = P.pool; c.cpool = P.cpool; c.ptab = P.ptab; c.fmsine = s_sine; c.f12tab = P.f12tab;
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
        // Drop-in mode: control writes the host made AFTER the voice's last segment of this
        // fragment (stamped with the frame where its next segment would start, i.e. up to 64) have
        // no boundary left to be applied at - the reference applies a write immediately
        // (a2_units.h:115), so they take effect here, before the state goes back to HBM.
        if (expl)
            while (evp < eve) {
                const uint4 e = P.ev[evp++];
                const int kind = e.y & 0xff, unit = (e.y >> 8) & 0xff, reg = (e.y >> 16) & 0xff;
                if (kind == EV_WRITE) ch.write(c, unit, reg, (int)e.z, (int)(e.x & 0xff), (int)e.w);
                else if (kind == EV_INIT) ch.init_unit(c, unit, (int)e.z, e.x & 0xff);
                else if (kind == EV_START) alive = 1;
                else if (kind == EV_STOP) alive = 0;
            }
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
    int general;            // host: a root ramp may be in flight - one CTA replays segments, no steady path
    int out_fmt;            // a2cu_set_output_format: 0 int32 8:24, 1 float32, 2 int16 (master block only)
};

// The driver edge, fused into the root stage: what the reference's audio drivers and wave writer do
// with the int32 8:24 master buffers - float = v * (1 / 8388608) (drivers/sdldrv.c:55-65, jackdrv
// alike), int16 = v >> 8 (a2_RenderWave -> a2_WaveWrite(A2_I24), waves.c:174-176). `i` indexes
// samples of the interleaved block [frame][channel].
A2CU_DEV void master_put(const MixParams &P, int i, int v) {
    if (P.out_fmt == 1) reinterpret_cast<float *>(P.master)[i] = __int2float_rn(v) * (1.0f / 8388608.0f);
    else if (P.out_fmt == 2) reinterpret_cast<short *>(P.master)[i] = (short)(v >> 8);
    else P.master[i] = v;
}

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
    // Steady state (no root events in the window, both rampers at rest): a2_PrepareRamper leaves
    // value = target, delta = 0 in EVERY segment (a2_dsp.h:130-134) - however the root's wake-ups
    // cut the window - so every frame uses the same two gains and the whole grid evaluates frames
    // independently. Otherwise CTA 0 replays the segments.
    // The host sets P.general (and launches ONE CTA) for every window a root write or ramp can
    // reach, so CTAs of one launch never disagree about `steady` while CTA 0 rewrites rstate.
    const int4 ra = __ldcg(reinterpret_cast<const int4 *>(P.rstate));           // vol: value target delta timer
    const int4 rb = __ldcg(reinterpret_cast<const int4 *>(P.rstate) + 1);       // pan
    const bool steady = !P.general && P.nev == 0 && ra.w == 0 && rb.w == 0 && ra.x == ra.y && rb.x == rb.y;
    if (steady) {
        const int v = ra.y, pn = rb.y;                          // targets
        const int vp = mulshr(pn, v, 24);
        int v0 = wsub(v, vp), v1 = wadd(v, vp);
        if (pn > 0xffffff || pn < -0xffffff) {
            const int lim = (int)((unsigned)v << 1);            // panmix.c:117-135 clamp variant
            if (v0 > lim) v0 = lim;
            if (v1 > lim) v1 = lim;
        }
        // latency-bound pass over a few KB: every thread requests its (up to four) frames first
        int2 *root2 = reinterpret_cast<int2 *>(root);
        for (int f0 = gtid; f0 < P.W; f0 += 4 * gsize) {
            int2 in[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) in[j] = f0 + j * gsize < P.W ? __ldcg(root2 + f0 + j * gsize) : make_int2(0, 0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int f = f0 + j * gsize;
                if (f >= P.W) break;
                if (clear) root2[f] = make_int2(0, 0);
                if (mono) master_put(P, f, (int)(((long long)in[j].x * v0 + (long long)in[j].y * v1) >> 25));
                else { master_put(P, f * 2, mulshr(in[j].x, v0, 24)); master_put(P, f * 2 + 1, mulshr(in[j].y, v1, 24)); }
            }
        }
        if (gtid == 0 && P.W > 0) { P.rstate[2] = 0; P.rstate[6] = 0; }     // deltas as PrepareRamper leaves them
        return;
    }
    if (!cta0) return;
    pm_bus(P, -1, P.rstate, root, mono, [&](int f, int r0, int r1) {
        if (mono) master_put(P, f, r0);
        else { master_put(P, f * 2, r0); master_put(P, f * 2 + 1, r1); }
        if (clear) { root[f * 2] = 0; root[f * 2 + 1] = 0; }
    });
}

// Root stage of the window BEFORE this launch (lagged exchange): sum its rows - published one
// launch ago by every rank - and run the root panmix into that window's output block.
// `rstate`, `channels`, `root_stage_on` are those of the engine (they do not change per window).
A2CU_DEV void xchg_finish_previous(const XchgParams &X, int *rstate, int channels, int root_stage_on, int tid,
                                   int nthreads, int out_fmt = 0) {
    xchg_collect(X, X.prev_epoch, X.sum, X.prev_W, tid, nthreads);
    MixParams M;
    M.acc = X.sum; M.W = X.prev_W; M.buffer = X.prev_buffer; M.ngroups = 0; M.channels = channels;
    M.nsplits = X.prev_nsplits;
    for (int i = 0; i < kMaxSplits; ++i) M.splits[i] = X.prev_splits[i];
    M.gstate = nullptr; M.rstate = rstate; M.ev = nullptr; M.nev = 0;
    M.master = X.prev_master; M.root_stage = root_stage_on; M.clear = 0; M.general = 0; M.out_fmt = out_fmt;
    root_stage(M, tid, nthreads, true);
    __syncthreads();
}

}  // namespace a2cu
