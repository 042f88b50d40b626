// a2cu_split.cuh - warp-specialised render kernel for wavetable voices.
//
// render_bank (a2cu_kernels.cuh) gives every voice one thread that walks the
// window frame by frame. With a few thousand voices that leaves most of the
// chip idle: 4 096 voices are 128 warps on 592 schedulers, each warp issuing a
// dependent instruction every ~4 cycles. But only the filter12 recurrence is
// truly serial in time. Inside one Process() segment
//
//   wtosc   sample k = Hermite(table, ph0 + k*dph) * (a0 + k*astep)   (closed form)
//   panmix  gains at k = vol0 + k*vstep, pan0 + k*pstep                (closed form)
//
// (a2_RunRamper and "ph += dph" are exact modular recurrences, wtosc.c:231-232,
// a2_dsp.h:152-155), so those two stages are evaluated per FRAME by helper
// warps while one warp per 32 voices runs the control-rate code and another the
// filter recurrence. One CTA = 32 voices (lane = voice everywhere, so the SoA
// state stays coalesced and the bus reduce stays a warp redux.sync):
//
//   control     (warp NH) events, per-segment prologues (the same unit code as
//               render_bank: WtOsc/Filter12/PanMix ::write/prepare/finish),
//               publishes per-segment parameters, advances state in closed form
//   serial      (warp NH+1) filter12 recurrence (filter12.c:97-118), frame by frame
//   NH helpers  stage A: oscillator samples -> tile A; lane = voice, frames sliced
//               stage C: panmix + bus sum <- tile A/B; lane = FRAME, each helper
//               sums a few voices for 32 frames into a shared-memory bus
//               (no cross-lane traffic); flushed with one coalesced atomic per
//               frame and channel one iteration later (stage D)
//
// The stages form a software pipeline over the window's fragments (ring of 4
// fragment slots in shared memory, one __syncthreads per fragment):
//   iteration i:  control(i)  stageA(i-1)  serial(i-2)  stageC(i-3)  flush(i-4)
//
// Eligibility is decided by the host per launch (a2cu_engine.cu): at most
// kSplitSegs segments per voice and fragment, only waves with a coefficient
// table, no noise / non-mipmapped waves in the bank. Otherwise render_bank runs
// on the same state layout. Results are bit-identical (tests).
#pragma once
#include "a2cu_kernels.cuh"

namespace a2cu {

constexpr int kRing = 4;
constexpr int kTileStride = 33;     // tile rows are voices; 33 keeps both access patterns conflict-free
// Warp roles: NH helper warps, one control warp, [one serial warp if FILT].
// Every helper runs a slice of stage A (lane = voice) and a part of stage C
// (lane = frame). Warp id % 4 is the SM sub-partition (one issue port each,
// B300_MICROARCH.md "Instruction issue"): the filter12 recurrence is one long
// dependent chain, and any other warp on its sub-partition takes issue slots
// (and holds them while it has independent work), so with FILT the serial
// warp gets sub-partition 3 to itself: it is the LAST warp (id % 4 == 3), the
// other warps with id % 4 == 3 stay idle, helpers and the control warp use
// ids with id % 4 != 3 in ascending order (logical index = id - id / 4).
template <bool FILT, int NH> struct SplitWarps {
    static constexpr int rows = (NH + 1 + 2) / 3;               // FILT: groups of 4 warp ids
    static constexpr int total = FILT ? rows * 4 : NH + 1;
    static constexpr int threads = total * 32;
    static constexpr int slice = (kMaxFrag + NH - 1) / NH;      // stage A: frames per helper
    static constexpr int groups = NH / 2;                       // stage C: voice groups (x 2 frame halves)
    static constexpr int vpg = (32 + groups - 1) / groups;      // voices per group
};

typedef WtOsc<true, false> SOsc;
typedef Filter12<1, false, false> SFilt;
typedef PanMix<1, 2, true, true> SPan;

template <int NOSC, bool FILT>
struct SplitLayout {
    // int offsets into dynamic shared memory
    static constexpr int oscp = 0;                                           // [ring][seg][osc][6][32]
    static constexpr int fp = oscp + kRing * kSplitSegs * NOSC * 6 * 32;     // [ring][seg][7][32]
    static constexpr int pmp = fp + (FILT ? kRing * kSplitSegs * 7 * 32 : 0);  // [ring][seg][5][32]
    static constexpr int split = pmp + kRing * kSplitSegs * 5 * 32;          // [ring][32]
    static constexpr int flags = split + kRing * 32;                         // [ring][32]
    static constexpr int meta = flags + kRing * 32;                          // [ring][2]: f0, n
    static constexpr int bus = meta + kRing * 2;                             // [32] bus of each voice
    static constexpr int sacc = bus + 32;                                    // [ring][64][2] home-bus sums
    static constexpr int smeta = sacc + kRing * kMaxFrag * 2;                // [ring][2]: f0, n for the flush
    static constexpr int tileA = smeta + kRing * 2;                          // [ring][64][33]
    static constexpr int tileB = tileA + kRing * kMaxFrag * kTileStride;     // [ring][64][33] (FILT)
    static constexpr int total = tileB + (FILT ? kRing * kMaxFrag * kTileStride : 0);
    static constexpr int table = (total + 31) & ~31;                         // staged coefficient table (int4)
    static constexpr size_t bytes = (size_t)table * sizeof(int);            // + 16 B per staged entry
    static constexpr int words = 1 + 14 * NOSC + (FILT ? 14 : 0) + 8;       // state words per voice
    static constexpr int filt_w = 1 + 14 * NOSC;                             // first word of filter12
    static constexpr int pm_w = filt_w + (FILT ? 14 : 0);
};

template <int NOSC, bool FILT, int NA>
__global__ void __launch_bounds__(SplitWarps<FILT, NA>::threads) render_split(const RenderParams P) {
    typedef SplitLayout<NOSC, FILT> L;
    typedef SplitWarps<FILT, NA> WR;
    constexpr int kSlice = WR::slice;
    extern __shared__ __align__(128) int sm[];
    __shared__ __align__(8) unsigned long long s_mbar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long t_entry = P.prof ? clock64() : 0;
    // logical role index (see SplitWarps): helpers 0..NA-1, control NA; -1 = idle
    const int hq = FILT ? ((warp & 3) == 3 ? -1 : warp - (warp >> 2)) : warp;
    const bool is_helper = hq >= 0 && hq < NA;
    const bool is_ctl = hq == NA;
    const bool is_ser = FILT && warp == WR::total - 1;
    const int v = blockIdx.x * 32 + lane;
    const bool valid = v < P.nvoices;
    const int W = P.W;

    int nfrag = 0;
    for (int f = 0; f < W; f = frag_end(f, P.buffer, W)) ++nfrag;
    const int mybus = valid ? P.bus_of[v] : -1;
    const int home = __shfl_sync(0xffffffffu, mybus, 0);
    if (is_ctl) sm[L::bus + lane] = mybus;
    for (int i = tid; i < kRing * kMaxFrag * 2; i += WR::threads) sm[L::sacc + i] = 0;
    // Stage the bank's wavetable (Hermite coefficient form, all mip levels) into
    // shared memory: one elected thread arms an mbarrier with the byte count and
    // issues TMA bulk copies; the helper warps wait on it before their first gather.
    const int4 *s_tab = reinterpret_cast<const int4 *>(sm + L::table);
    const int stage_n = P.stage_count;
    if (stage_n) {
        if (tid == 0) mbar_init(&s_mbar, 1);
        __syncthreads();
        if (tid == 0) {
            const unsigned total_b = (unsigned)stage_n * 16u;
            mbar_expect_tx(&s_mbar, total_b);
            const char *src = reinterpret_cast<const char *>(P.cpool + P.stage_begin);
            char *dst = reinterpret_cast<char *>(sm + L::table);
            for (unsigned off = 0; off < total_b; off += 32768u)
                tma_bulk_g2s(dst + off, src + off, min(32768u, total_b - off), &s_mbar);
        }
    }
    bool tab_ready = stage_n == 0;

    Ctx c;
    c.waves = P.waves; c.pool = P.pool; c.cpool = P.cpool; c.ptab = P.ptab; c.fmsine = nullptr; c.f12tab = P.f12tab;
    c.samplerate = P.samplerate;
    StatePtr sp{P.state + (valid ? v : 0), P.stride};

    // ---- control warp state ----
    SOsc osc[NOSC];
    SFilt filt;
    SPan pm;
    int alive = 0;
    unsigned evp = 0, eve = 0;
    int next_ev = 0x7fffffff, seg_end = 0;
    bool in_seg = false;
    int cf0 = 0;
    // ---- serial warp state ----
    int d1 = 0, d2 = 0;

    if (is_ctl && valid) {
        alive = sp.ld(0) & 1;
#pragma unroll
        for (int i = 0; i < NOSC; ++i) osc[i].load(sp, 1 + 14 * i);
        if (FILT) filt.load(sp, L::filt_w);
        pm.load(sp, L::pm_w);
        if (P.ev_off) { evp = P.ev_off[v]; eve = P.ev_off[v + 1]; }
        next_ev = evp < eve ? (int)(P.ev[evp].x >> 8) : 0x7fffffff;
    }
    if (is_ser && valid) {
        d1 = sp.ld(L::filt_w + 12);
        d2 = sp.ld(L::filt_w + 13);
    }

    const int lag_c = FILT ? 3 : 2;     // stage C runs this many iterations behind control
    const long long t_loop = P.prof ? clock64() : 0;
    for (int it = 0; it < nfrag + lag_c + 1; ++it) {
        long long t_begin = 0, t_mid = 0;
        if (P.prof) t_begin = clock64();
        // ================= control(it) =================
        if (is_ctl && it < nfrag) {
            const int slot = it % kRing;
            const int f0 = cf0;
            const int fe = frag_end(f0, P.buffer, W);
            int split = fe - f0, flags = 0;
            if (valid) {
                int f = f0, seg = 0;
                while (f < fe) {
                    // ---- segment boundary: identical to render_bank ----
                    if (in_seg) {
#pragma unroll
                        for (int i = 0; i < NOSC; ++i) osc[i].finish();
                        if (FILT) filt.finish();
                        pm.finish();
                    }
                    while (next_ev <= f) {
                        const uint4 e = P.ev[evp];
                        const int kind = e.y & 0xff, unit = (e.y >> 8) & 0xff, reg = (e.y >> 16) & 0xff;
                        const int st = (int)(e.x & 0xff);
                        if (kind == EV_WRITE || kind == EV_INIT) {
                            const bool init = kind == EV_INIT;
#pragma unroll
                            for (int i = 0; i < NOSC; ++i)
                                if (unit == i) {
                                    if (init) osc[i].init(c, (int)e.z, (unsigned)st);
                                    else osc[i].write(c, reg, (int)e.z, st, (int)e.w);
                                }
                            if (FILT && unit == NOSC) {
                                if (init) filt.init(c, (int)e.z, (unsigned)st);
                                else filt.write(c, reg, (int)e.z, st, (int)e.w);
                            }
                            if (unit == NOSC + (FILT ? 1 : 0)) {
                                if (init) pm.init(c, (int)e.z, (unsigned)st);
                                else pm.write(c, reg, (int)e.z, st, (int)e.w);
                            }
                        } else if (kind == EV_START) alive = 1;
                        else if (kind == EV_STOP) alive = 0;
                        ++evp;
                        next_ev = evp < eve ? (int)(P.ev[evp].x >> 8) : 0x7fffffff;
                    }
                    int nxt = min(fe, next_ev);
                    for (int k = 0; k < P.nsplits; ++k)
                        if (P.splits[k] > f) nxt = min(nxt, P.splits[k]);
                    seg_end = nxt;
                    in_seg = alive != 0;
                    const int n = nxt - f;
                    if (in_seg) {
#pragma unroll
                        for (int i = 0; i < NOSC; ++i) osc[i].prepare(c, n);
                        if (FILT) filt.prepare(c, n);
                        pm.prepare(c, n);
                    }
                    // ---- publish the segment, advance in closed form ----
                    if (seg < kSplitSegs) {
                        if (in_seg) flags |= 1 << seg;
                        const int sb = (slot * kSplitSegs + seg);
#pragma unroll
                        for (int i = 0; i < NOSC; ++i) {
                            int *q = sm + L::oscp + ((sb * NOSC + i) * 6) * 32 + lane;
                            const bool live = in_seg && osc[i].run == RUN_TABLE && osc[i].cf;
                            q[0] = live ? (int)(osc[i].cf - c.cpool) : -1;
                            q[32] = (int)(unsigned)osc[i].ph;
                            q[64] = (int)(unsigned)(osc[i].ph >> 32);
                            q[96] = (int)osc[i].dph;
                            q[128] = osc[i].a.value;
                            q[160] = osc[i].astep;
                            if (live) {     // wtosc.c:231-232 over n frames
                                osc[i].ph += (unsigned long long)osc[i].dph * (unsigned)n;
                                osc[i].a.value = wadd(osc[i].a.value, wmul(osc[i].astep, n));
                            }
                        }
                        if (FILT) {
                            int *q = sm + L::fp + (sb * 7) * 32 + lane;
                            q[0] = filt.f0; q[32] = filt.df; q[64] = filt.q.value; q[96] = filt.qstep;
                            q[128] = filt.lp; q[160] = filt.bp; q[192] = filt.hp;
                            if (in_seg) filt.q.value = wadd(filt.q.value, wmul(filt.qstep, n));
                        }
                        {
                            int *q = sm + L::pmp + (sb * 5) * 32 + lane;
                            q[0] = pm.vol.value; q[32] = pm.vstep; q[64] = pm.pan.value; q[96] = pm.pstep;
                            q[128] = pm.clamp ? 1 : 0;
                            if (in_seg) {
                                pm.vol.value = wadd(pm.vol.value, wmul(pm.vstep, n));
                                pm.pan.value = wadd(pm.pan.value, wmul(pm.pstep, n));
                            }
                        }
                        if (seg == 0) split = nxt - f0;
                    }
                    f = nxt;
                    ++seg;
                }
            }
            sm[L::split + slot * 32 + lane] = split;
            sm[L::flags + slot * 32 + lane] = flags;
            if (lane == 0) { sm[L::meta + slot * 2] = f0; sm[L::meta + slot * 2 + 1] = fe - f0; }
            cf0 = fe;
        }
        // ================= serial(it - 2): filter12 recurrence =================
        if (is_ser && it >= 2 && it - 2 < nfrag) {
            const int slot = (it - 2) % kRing;
            const int n = sm[L::meta + slot * 2 + 1];
            const int split = sm[L::split + slot * 32 + lane];
            const int flags = sm[L::flags + slot * 32 + lane];
            const int *ta = sm + L::tileA + slot * kMaxFrag * kTileStride + lane;
            int *tb = sm + L::tileB + slot * kMaxFrag * kTileStride + lane;
            for (int seg = 0; seg < kSplitSegs; ++seg) {
                const int a = seg ? split : 0, b = seg ? n : min(split, n);
                if (a >= b) continue;
                if (!((flags >> seg) & 1)) continue;    // inactive: state frozen, output unused
                const int *q = sm + L::fp + ((slot * kSplitSegs + seg) * 7) * 32 + lane;
                int f0v = q[0];
                const int df = q[32];
                int qv = q[64];
                const int qstep = q[96], lp = q[128], bp = q[160], hp = q[192];
#pragma unroll 4
                for (int f = a; f < b; ++f) {           // filter12.c:97-118
                    const int fc = f0v >> 12, qq = qv >> 12;
                    const int in = ta[f * kTileStride];
                    const int d1s = d1 >> 4;
                    const int l = wadd(d2, wmul(fc, d1s) >> 8);
                    // (in>>5) - l - (q*d1s>>8). ptxas turns this into two shift-adds after l
                    // (LEA.HI.SX32); forcing "t = (in>>5) - q-term off the chain, h = t - l" with an
                    // inline PTX sub puts an IMAD.IADD on the other pipe and measured 5 % SLOWER.
                    const int h = wsub(wsub(in >> 5, wmul(qq, d1s) >> 8), l);
                    const int bb = wadd(wmul(fc, h >> 4) >> 8, d1);
                    tb[f * kTileStride] = wadd(wadd(wmul(l, lp), wmul(bb, bp)), wmul(h, hp)) >> 3;
                    d1 = bb; d2 = l;
                    f0v = wadd(f0v, df);
                    qv = wadd(qv, qstep);
                }
            }
        }
        // ================= stage A(it - 1): oscillators, lane = voice, frames sliced =================
        if (is_helper && it >= 1 && it - 1 < nfrag) {
            const int h = hq;
            const int slot = (it - 1) % kRing;
            const int n = sm[L::meta + slot * 2 + 1];
            const int split = sm[L::split + slot * 32 + lane];
            const int flags = sm[L::flags + slot * 32 + lane];
            int *ta = sm + L::tileA + slot * kMaxFrag * kTileStride + lane;
            const int s0 = h * kSlice, s1 = min(n, s0 + kSlice);
            for (int seg = 0; seg < kSplitSegs; ++seg) {
                const int sa = seg ? split : 0;
                const int a = max(s0, sa), b = min(s1, seg ? n : split);
                if (a >= b) continue;
                int acc[kSlice];
#pragma unroll
                for (int k = 0; k < kSlice; ++k) acc[k] = 0;
                if ((flags >> seg) & 1) {
#pragma unroll 1
                    for (int i = 0; i < NOSC; ++i) {    // not unrolled: keeps the hot loop in the I-cache
                        const int *q = sm + L::oscp + (((slot * kSplitSegs + seg) * NOSC + i) * 6) * 32 + lane;
                        const int cfo = q[0];
                        if (cfo < 0) continue;          // silent segment of this oscillator
                        if (!tab_ready) { mbar_wait(&s_mbar, 0); tab_ready = true; }
                        const int srel = cfo - P.stage_begin;
                        // generic pointer: the staged copy in shared memory or the pool in global memory
                        const int4 *cf = (srel >= 0 && srel < stage_n - 64) ? s_tab + srel : c.cpool + cfo;
                        const unsigned dph = (unsigned)q[96];
                        unsigned long long ph = ((unsigned long long)(unsigned)q[64] << 32) | (unsigned)q[32];
                        ph += (unsigned long long)dph * (unsigned)(a - sa);
                        const int astep = q[160];
                        const int av0 = wadd(q[128], wmul(astep, a - sa));
                        const unsigned half = dph >> 17;
                        // all kSlice frames are evaluated (independent chains the
                        // scheduler can overlap); frames past the segment end read
                        // table slack and are dropped at the store (wtosc.c:226-233)
#pragma unroll
                        for (int k = 0; k < kSlice; ++k) {
                            const unsigned p16 = (unsigned)((ph + (unsigned long long)dph * (unsigned)k) >> 16);
                            const int hv = hermite_cf_smem(cf, p16) + hermite_cf_smem(cf, p16 + half);
                            acc[k] = wadd(acc[k], mulshr(hv, wadd(av0, wmul(astep, k)), 17));
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < kSlice; ++k)
                    if (a + k < b) ta[(a + k) * kTileStride] = acc[k];
            }
        }
        if (P.prof) t_mid = clock64();
        // ================= stage C(it - lag_c): panmix + bus sum, lane = frame =================
        if (is_helper && hq < 2 * WR::groups && it >= lag_c && it - lag_c < nfrag) {
            const int h = hq;
            const int slot = (it - lag_c) % kRing;
            const int f0 = sm[L::meta + slot * 2];
            const int n = sm[L::meta + slot * 2 + 1];
            const int f = (h & 1) * 32 + lane;              // my frame of the fragment
            const int vlo = (h >> 1) * WR::vpg;
            if (h == 0 && lane == 0) {                      // the params slot is recycled before the flush
                sm[L::smeta + slot * 2] = f0;
                sm[L::smeta + slot * 2 + 1] = n;
            }
            const int *tin = sm + (FILT ? L::tileB : L::tileA) + slot * kMaxFrag * kTileStride + f * kTileStride;
            int sum0 = 0, sum1 = 0;
            if (f < n) {
#pragma unroll
                for (int j = 0; j < WR::vpg; ++j) {
                    const int vv = vlo + j;
                    if (vv >= 32) break;
                    const int split = sm[L::split + slot * 32 + vv];
                    const int flags = sm[L::flags + slot * 32 + vv];
                    const int seg = f >= split ? 1 : 0;
                    if (!((flags >> seg) & 1)) continue;
                    const int *q = sm + L::pmp + ((slot * kSplitSegs + seg) * 5) * 32 + vv;
                    const int k = f - (seg ? split : 0);    // panmix.c:78-115
                    const int vol = wadd(q[0], wmul(q[32], k));
                    const int pan = wadd(q[64], wmul(q[96], k));
                    const int vp = mulshr(pan, vol, 24);
                    int v0 = wsub(vol, vp), v1 = wadd(vol, vp);
                    if (q[128]) {
                        const int lim = (int)((unsigned)vol << 1);
                        if (v0 > lim) v0 = lim;
                        if (v1 > lim) v1 = lim;
                    }
                    const int in = tin[vv];
                    const int o0 = mulshr(in, v0, 24), o1 = mulshr(in, v1, 24);
                    const int vb = sm[L::bus + vv];
                    if (vb == home) { sum0 = wadd(sum0, o0); sum1 = wadd(sum1, o1); }
                    else {
                        int *a = P.acc + ((size_t)vb * W + f0 + f) * 2;
                        atomicAdd(a, o0);
                        atomicAdd(a + 1, o1);
                    }
                }
                int *sa = sm + L::sacc + (slot * kMaxFrag + f) * 2;
                if (sum0) atomicAdd(sa, sum0);
                if (sum1) atomicAdd(sa + 1, sum1);
            }
        }
        // ================= stage D(it - lag_c - 1): flush the fragment's bus sums =================
        if (hq == 0 && it >= lag_c + 1 && it - lag_c - 1 < nfrag && home >= 0) {
            const int slot = (it - lag_c - 1) % kRing;
            const int f0 = sm[L::smeta + slot * 2];
            const int n = sm[L::smeta + slot * 2 + 1];
            int *sa = sm + L::sacc + slot * kMaxFrag * 2;
            int *ga = P.acc + ((size_t)home * W + f0) * 2;
            for (int i = lane; i < n * 2; i += 32) {
                const int val = sa[i];
                if (val) { atomicAdd(ga + i, val); sa[i] = 0; }
            }
        }
        long long t_done = 0;
        if (P.prof) t_done = clock64();
        __syncthreads();
        if (P.prof && lane == 0) {
            // [0] control [1] serial [2] stage A [3] stage C [4] barrier wait, [5] iterations
            if (is_ctl) atomicAdd(P.prof + 0, (unsigned long long)(t_done - t_begin));
            else if (is_ser) atomicAdd(P.prof + 1, (unsigned long long)(t_done - t_begin));
            else if (hq == 1) {
                atomicAdd(P.prof + 2, (unsigned long long)(t_mid - t_begin));
                atomicAdd(P.prof + 3, (unsigned long long)(t_done - t_mid));
                atomicAdd(P.prof + 4, (unsigned long long)(clock64() - t_done));
                atomicAdd(P.prof + 5, 1ull);
            }
        }
    }

    if (is_ctl && valid) {
        if (in_seg) {
#pragma unroll
            for (int i = 0; i < NOSC; ++i) osc[i].finish();
            if (FILT) filt.finish();
            pm.finish();
        }
        sp.st(0, alive);
#pragma unroll
        for (int i = 0; i < NOSC; ++i) osc[i].store(sp, 1 + 14 * i);
        if (FILT) filt.store(sp, L::filt_w);
        pm.store(sp, L::pm_w);
    }
    __syncthreads();
    if (is_ser && valid) {       // the recurrence state lives in the serial warp
        sp.st(L::filt_w + 12, d1);
        sp.st(L::filt_w + 13, d2);
    }
    if (P.prof && tid == 0) {       // [6] prologue, [7] loop, per CTA (a2cu_split_profile)
        atomicAdd(P.prof + 6, (unsigned long long)(t_loop - t_entry));
        atomicAdd(P.prof + 7, (unsigned long long)(clock64() - t_loop));
    }
    // ---- fused root stage: the last CTA to get here owns the finished root bus ----
    if (P.fuse_root) {
        __shared__ int s_last;
        __syncthreads();                    // every bus flush of this CTA has been issued
        if (tid == 0) {
            __threadfence();                // ... and is visible before the ticket is taken
            s_last = atomicAdd(P.fuse_counter, 1u) == gridDim.x - 1 ? 1 : 0;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            MixParams M;
            M.acc = P.acc; M.W = P.W; M.buffer = P.buffer; M.ngroups = 0; M.channels = P.fuse_channels;
            M.nsplits = P.nsplits;          // root wake-ups cut the root panmix's segments too
            for (int i = 0; i < kMaxSplits; ++i) M.splits[i] = P.splits[i];
            M.gstate = nullptr; M.rstate = P.fuse_rstate; M.ev = nullptr; M.nev = 0;
            M.master = P.fuse_master; M.root_stage = P.fuse_root_stage; M.clear = 1; M.general = 0;
            // sharded render: the root bus of all ranks is summed here, through NVLink peer memory
            if (P.xchg.world > 1) xchg_root_bus(P.xchg, P.acc, P.W, tid, WR::threads);
            root_stage(M, tid, WR::threads, true);
            if (tid == 0) *P.fuse_counter = 0u;     // ready for the next launch
        }
    }
}

}  // namespace a2cu
