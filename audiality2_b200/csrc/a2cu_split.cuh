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
// filter recurrence. A CTA renders VS "voice sets" of 32 voices; lane = voice in
// the control, recurrence and oscillator roles (the SoA state stays coalesced),
// lane = frame in the panmix / bus role. Roles of one set:
//
//   control     events, per-segment prologues (the same unit code as render_bank:
//               WtOsc/Filter12/PanMix ::write/prepare/finish), publishes the
//               per-segment parameters, advances state in closed form
//   serial      filter12 recurrence (filter12.c:97-118), frame by frame, in place
//               on the fragment's tile
//   NH helpers  stage A: oscillator samples -> tile (lane = voice, frames sliced);
//               taps come from the Hermite-coefficient table staged in shared
//               memory by TMA, or - large sampled waves - straight from the int16
//               pool in HBM (closed-form phase: every helper has 2 x slice
//               independent sector requests in flight per lane)
//               stage C: panmix + bus sum <- tile (lane = FRAME; each helper sums
//               a few voices for 32 frames and adds them to the device bus with
//               one coalesced reduction per channel)
//
// The roles form a pipeline over the window's fragments through a ring of R
// fragment slots in shared memory. Every edge is an mbarrier per slot
// (producer arrives, consumer waits on the slot's phase parity):
//
//   control --P--> helpers(A) --A--> serial --B--> helpers(C) --C--> control (slot free)
//
// so the recurrence warp - the critical path of a filtered voice - runs back to
// back, one fragment after the other, and nothing ever waits for the slowest
// warp of the CTA (round 1 used one __syncthreads per fragment: 4 of 19
// iterations were pipeline fill/drain and every role ran at the pace of the
// slowest).
//
// Eligibility is decided by the host per launch (a2cu_engine.cu): at most
// kSplitSegs segments per voice and fragment, no noise / one-shot sampled waves
// in the bank. Otherwise render_bank runs on the same state layout. Results are
// bit-identical (tests).
#pragma once
#include "a2cu_kernels.cuh"

namespace a2cu {

#ifndef A2CU_SER_PF
#define A2CU_SER_PF 16
#endif
// Stage A work split. 0: lane = voice, the fragment's frames sliced over the helpers. 1: lane = FRAME,
// one voice (and 32-frame half) at a time: the 64 taps of a warp instruction then lie on a short
// contiguous run of the wave - DRAM pages and cache lines are read whole instead of a sector here and
// there (the HBM-bound gather), and a shared-memory table gather meets fewer bank conflicts.
#ifndef A2CU_LF_RAW
#define A2CU_LF_RAW 1
#endif
#ifndef A2CU_LF_BATCH
#define A2CU_LF_BATCH 1
#endif
#ifndef A2CU_LF_TABLE
#define A2CU_LF_TABLE 0
#endif
constexpr int kTileStride = 33;     // tile rows are frames, columns voices; 33 keeps both access patterns conflict-free
constexpr int kOscWords = 8;        // published per oscillator and segment

// Warp roles. Warp id % 4 is the SM sub-partition (one issue port each,
// B300_MICROARCH.md "Instruction issue"). The filter12 recurrence is one long
// dependent chain: any other warp on its sub-partition takes issue slots, so
// with FILT the serial warps own sub-partition 3 (warp ids 3, 7, ...): the
// other warps with id % 4 == 3 stay idle, helpers and control warps use the ids
// with id % 4 != 3 in ascending order (logical index q = id - id / 4).
// q < VS * NH: helper (set q / NH, index q % NH); then VS control warps.
// (Measured: helpers on sub-partition 3 cost the recurrence warp 25-40 % of its speed even when it
// has the highest warp id, i.e. issue priority - profiles/README.md - so that port stays its own.
// What still slows it in the kernel - 60 cycles per frame against 45 alone - is shared-memory
// contention with the helpers' gathers: profiles/ubench_serial.cu reproduces 73 / 127 cycles per
// frame with 6 / 11 warps gathering. An 8-byte tap table instead of the 16-byte coefficient entries
// halves the gather wavefronts and brought the chain to 52 cycles, but the extra arithmetic in
// stage A and in render_bank cost more than that gained: 65.7 vs 68.7 G on cfg2, 25.8 vs 29.5 G on
// cfg3, 161 vs 173 G at 262 144 voices - measured and reverted.)
template <bool FILT, int NH, int VS> struct SplitWarps {
    static constexpr int workers = VS * (NH + 1);
    static constexpr int rows = (workers + 2) / 3 > VS ? (workers + 2) / 3 : VS;     // FILT: groups of 4 warp ids
    static constexpr int total = FILT ? rows * 4 : workers;
    static constexpr int threads = total * 32;
    static constexpr int slice = (kMaxFrag + NH - 1) / NH;      // stage A: frames per helper
    static constexpr int groups = NH / 2;                       // stage C: voice groups (x 2 frame halves)
    static constexpr int vpg = (32 + groups - 1) / groups;      // voices per group
};

typedef WtOsc<true, false> SOsc;
typedef Filter12<1, false, false> SFilt;
typedef PanMix<1, 2, true, true> SPan;

template <int NOSC, bool FILT, int R>
struct SplitLayout {
    // int offsets into one voice set's part of dynamic shared memory
    static constexpr int oscp = 0;                                               // [R][seg][osc][8][32]
    static constexpr int fp = oscp + R * kSplitSegs * NOSC * kOscWords * 32;     // [R][seg][7][32]
    static constexpr int pmp = fp + (FILT ? R * kSplitSegs * 7 * 32 : 0);        // [R][seg][8][32]
    static constexpr int split = pmp + R * kSplitSegs * 8 * 32;                  // [R][32]
    static constexpr int flags = split + R * 32;                                 // [R][32]
    static constexpr int meta = flags + R * 32;                                  // [R][2]: f0, n
    static constexpr int bus = meta + R * 2;                                     // [32] bus of each voice
    static constexpr int tile = ((bus + 32) + 31) & ~31;                         // [R][64][33]
    static constexpr int set_ints = ((tile + R * kMaxFrag * kTileStride) + 31) & ~31;
    static constexpr int words = 1 + 14 * NOSC + (FILT ? 14 : 0) + 8;           // state words per voice
    static constexpr int filt_w = 1 + 14 * NOSC;                                 // first word of filter12
    static constexpr int pm_w = filt_w + (FILT ? 14 : 0);
};
// Each voice set's contribution to its home bus is accumulated in shared memory over the WHOLE
// window ([kSplitMaxWin][2] ints per set) and goes to the device bus once, at the end of the launch. The root bus
// is a handful of cache lines that every CTA of the grid adds into and L2 serialises same-line
// atomics (~20 cycles per 128-byte request): flushing per fragment made every CTA wait ~2.8 k
// cycles per fragment for the other 127 (profiles/r02_split_timeline_*.txt).
template <int NOSC, bool FILT, int R, int VS>
constexpr size_t split_smem_bytes() {     // + 16 B per staged table entry
    return (size_t)VS * (SplitLayout<NOSC, FILT, R>::set_ints + kSplitMaxWin * 2) * sizeof(int);
}

A2CU_DEV void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// x % (wsize << 24) for a phase accumulator (24 fraction bits) less than 2^56: the fraction passes
// through, the integer part is a 32-bit modulo (a 64-bit one is ~130 instructions, which made this
// line a third of the gather kernel's instruction stream)
A2CU_DEV unsigned long long mod_wave(unsigned long long x, unsigned wsize) {
    const unsigned xi = (unsigned)(x >> 24);
    return ((unsigned long long)(xi % wsize) << 24) | (x & 0xffffffull);
}

// filter12.c:97-118 over frames [a, b) of one voice, in place on its tile column. Inputs are
// fetched A2CU_SER_PF frames ahead of the dependent chain: the helpers of a set all start their
// gathers on the same barrier, a burst of ~1.2 k shared-memory wavefronts that a load issued by
// this warp queues behind - the look-ahead has to cover it. (Tried and measured slower on the B200:
// volatile loads/stores to pin the prefetch before the chain, 4.6 k instead of 3.7 k cycles per
// fragment; taking (in>>5) - (q*d1s>>8) off the chain - ptxas already schedules the two shift-adds
// behind l well, the explicit form only adds an instruction.)
template <bool RAMP>
A2CU_DEV void filter_run(int *t, int a, int b, int &d1, int &d2, int f0v, int df, int qv, int qstep,
                         int lp, int bp, int hp) {
    int fc = f0v >> 12, qq = qv >> 12;
    auto step = [&](int in) -> int {
        const int d1s = d1 >> 4;
        const int l = wadd(d2, wmul(fc, d1s) >> 8);
        const int h = wsub(wsub(in >> 5, wmul(qq, d1s) >> 8), l);
        const int bb = wadd(wmul(fc, h >> 4) >> 8, d1);
        const int out = wadd(wadd(wmul(l, lp), wmul(bb, bp)), wmul(h, hp)) >> 3;
        d1 = bb; d2 = l;
        if (RAMP) {
            f0v = wadd(f0v, df); qv = wadd(qv, qstep);
            fc = f0v >> 12; qq = qv >> 12;
        }
        return out;
    };
    int f = a;
    constexpr int PF = A2CU_SER_PF;
    if (f + PF <= b) {
        int cur[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) cur[k] = t[(f + k) * kTileStride];
        while (true) {
            const bool more = f + 2 * PF <= b;
            int nxt[PF];
#pragma unroll
            for (int k = 0; k < PF; ++k) nxt[k] = more ? t[(f + PF + k) * kTileStride] : 0;
#pragma unroll
            for (int k = 0; k < PF; ++k) t[(f + k) * kTileStride] = step(cur[k]);
            f += PF;
            if (!more) break;
#pragma unroll
            for (int k = 0; k < PF; ++k) cur[k] = nxt[k];
        }
    }
    for (; f < b; ++f) t[f * kTileStride] = step(t[f * kTileStride]);
}

// RAW: the bank may play waves without a Hermite-coefficient table (large sampled waves): stage A
// then carries the raw-tap gather as well. Kept out of the table-only instantiation, whose hot loop
// should stay small (instruction-cache misses show up as `no_instruction` stalls in ncu).
template <int NOSC, bool FILT, int NH, int VS, int R, bool RAW>
__global__ void __launch_bounds__(SplitWarps<FILT, NH, VS>::threads, (RAW && NOSC == 1 && !FILT) ? 2 : 1)
render_split(const RenderParams P) {
    typedef SplitLayout<NOSC, FILT, R> L;
    typedef SplitWarps<FILT, NH, VS> WR;
    constexpr int kSlice = WR::slice;
    constexpr bool LF = RAW ? (A2CU_LF_RAW != 0) : (A2CU_LF_TABLE != 0);
    extern __shared__ __align__(128) int sm_all[];
    __shared__ __align__(8) unsigned long long s_mbar;               // table staging (TMA)
    __shared__ __align__(8) unsigned long long s_bar[VS][4][R];     // P, A, B, C per slot
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long t_entry = P.prof ? clock64() : 0;
    // wall-clock marks of the whole grid (globaltimer, ns) for a2cu_split_trace: role 4, fragments
    // 56..59 = first CTA entry (min), last pipeline end (max), last CTA exit before the tail (max),
    // end of the fused tail
    auto gmark = [&](int k, bool is_min) {
        if (P.prof && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            unsigned long long *p = P.prof + 8 + ((4 * 64 + 56 + k) * 2);
            if (is_min) atomicMin(p, t); else atomicMax(p, t);
        }
    };
    gmark(0, true);
    // ---- role of this warp ----
    const int q = FILT ? ((warp & 3) == 3 ? -1 : warp - (warp >> 2)) : warp;
    const bool is_helper = q >= 0 && q < VS * NH;
    const bool is_ctl = q >= VS * NH && q < VS * (NH + 1);
    const bool is_ser = FILT && (warp & 3) == 3 && (warp >> 2) < VS;
    const int set = is_helper ? q / NH : is_ctl ? q - VS * NH : is_ser ? (warp >> 2) : 0;
    const int hq = is_helper ? q % NH : -1;
    int *sm = sm_all + set * L::set_ints;
    unsigned long long *barP = s_bar[set][0], *barA = s_bar[set][1], *barB = s_bar[set][2], *barC = s_bar[set][3];
    const int v = (blockIdx.x * VS + set) * 32 + lane;
    const bool valid = v < P.nvoices;
    const int W = P.W;
    // The voice state comes from HBM (a different bank every window in a large project): the control
    // and recurrence warps request it first thing, so that the round trip overlaps the CTA prologue
    // (accumulator clear, barrier setup, table staging) instead of following it.
    StatePtr sp{P.state + (valid ? v : 0), P.stride};
    SOsc osc[NOSC];
    SFilt filt;
    SPan pm;
    int alive = 0;
    unsigned evp = 0, eve = 0;
    int next_ev = 0x7fffffff;
    int d1 = 0, d2 = 0;
    if (is_ctl && valid) {
        alive = sp.ld(0) & 1;
#pragma unroll
        for (int i = 0; i < NOSC; ++i) osc[i].load(sp, 1 + 14 * i);
        if (FILT) filt.load(sp, SplitLayout<NOSC, FILT, R>::filt_w);
        pm.load(sp, SplitLayout<NOSC, FILT, R>::pm_w);
        if (P.ev_off) { evp = P.ev_off[v]; eve = P.ev_off[v + 1]; }
        next_ev = evp < eve ? (int)(P.ev[evp].x >> 8) : 0x7fffffff;
    }
    if (FILT && is_ser && valid) {
        d1 = sp.ld(SplitLayout<NOSC, FILT, R>::filt_w + 12);
        d2 = sp.ld(SplitLayout<NOSC, FILT, R>::filt_w + 13);
    }

    // fragments restart at every driver buffer (frag_end): closed form instead of walking the window
    const int nfrag = (W / P.buffer) * ((P.buffer + kMaxFrag - 1) / kMaxFrag) +
                      ((W % P.buffer) + kMaxFrag - 1) / kMaxFrag;
    const int mybus = valid ? P.bus_of[v] : -1;
    // home bus of a voice set: the bus of its first voice
    const int home = (blockIdx.x * VS + set) * 32 < P.nvoices ? P.bus_of[(blockIdx.x * VS + set) * 32] : -1;
    if (is_ctl) sm[L::bus + lane] = mybus;
    int *wacc_all = sm_all + VS * L::set_ints;              // [VS][kSplitMaxWin][2] home-bus sums
    for (int i = tid; i < VS * kSplitMaxWin * 2; i += WR::threads) wacc_all[i] = 0;
    int *wacc = wacc_all + set * kSplitMaxWin * 2;
    // Stage the bank's wavetable (Hermite coefficient form, all mip levels) into shared memory:
    // one elected thread arms an mbarrier with the byte count and issues TMA bulk copies; the
    // helper warps wait on it before their first gather.
    int *sm_table = wacc_all + VS * kSplitMaxWin * 2;
    const int4 *s_tab = reinterpret_cast<const int4 *>(sm_table);
    const int stage_n = P.stage_count;
    if (tid == 0) {
        if (stage_n) mbar_init(&s_mbar, 1);
        for (int s = 0; s < VS; ++s)
            for (int k = 0; k < R; ++k) {
                mbar_init(&s_bar[s][0][k], 1);          // P: control
                mbar_init(&s_bar[s][1][k], NH);         // A: every helper
                mbar_init(&s_bar[s][2][k], 1);          // B: serial
                mbar_init(&s_bar[s][3][k], NH);         // C: every helper
            }
    }
    __syncthreads();
    if (stage_n && tid == 0) {
        const unsigned total_b = (unsigned)stage_n * 16u;
        mbar_expect_tx(&s_mbar, total_b);
        const char *src = reinterpret_cast<const char *>(P.cpool + P.stage_begin);
        char *dst = reinterpret_cast<char *>(sm_table);
        for (unsigned off = 0; off < total_b; off += 32768u)
            tma_bulk_g2s(dst + off, src + off, min(32768u, total_b - off), &s_mbar);
    }
    bool tab_ready = stage_n == 0;

    Ctx c;
    c.waves = P.waves; c.pool = P.pool; c.cpool = P.cpool; c.ptab = P.ptab; c.fmsine = nullptr; c.f12tab = P.f12tab;
    c.samplerate = P.samplerate;
    const long long t_loop = P.prof ? clock64() : 0;
    long long busy = 0, busy2 = 0;
    // timeline of CTA 0 / set 0 (a2cu_split_trace): prof[8 + ((role * 64 + fragment) * 2 + end)]
    // roles: 0 control, 1 serial, 2 / 3 stage A / C of helper 0, 4 / 5 of the last helper
    auto trace = [&](int role, int frag, int end) {
        if (P.prof && blockIdx.x == 0 && set == 0 && lane == 0 && frag < 64)
            P.prof[8 + ((role * 64 + frag) * 2 + end)] = (unsigned long long)(clock64() - t_loop);
    };

    // =====================================================================================
    // control: events, prologues, publish
    // =====================================================================================
    if (is_ctl) {
        bool in_seg = false;
        unsigned open_mask = 0;
        int f0 = 0;
        for (int it = 0; it < nfrag; ++it) {
            const int slot = it % R;
            if (it >= R) mbar_wait(&barC[slot], ((it / R) - 1) & 1);       // slot free again
            const long long tb = P.prof ? clock64() : 0;
            trace(0, it, 0);
            const int fe = frag_end(f0, P.buffer, W);
            int split = fe - f0, flags = 0;
            if (valid) {
                int f = f0, seg = 0;
                while (f < fe) {
                    // ---- segment boundary: as render_bank, except that an oscillator's
                    // finish() / prepare() pair is skipped while nothing can change (see below) ----
                    if (in_seg) {
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
                                    // an amplitude write (wtosc.c:498-501 only sets the ramper) leaves the
                                    // segment open: the fast path below redoes a2_PrepareRamper for it
                                    if (((open_mask >> i) & 1) && (init || reg != 2)) {
                                        osc[i].finish();
                                        open_mask &= ~(1u << i);
                                    }
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
                    in_seg = alive != 0;
                    const int n = nxt - f;
                    // Oscillators. `open_mask` bit i: oscillator i has a segment open (its finish() is
                    // still due). A mip-mapped oscillator whose PITCH ramper is at rest and that no event
                    // touched runs the same segment again: wtosc.c:239-286 would recompute the same
                    // mip level, increment and table, so only the amplitude ramper and the loop wrap /
                    // end test of the prologue are redone (phase >> mm << mm is the identity here: the
                    // low mm bits are already zero). This prologue is ~400 dependent instructions per
                    // oscillator; with eight oscillators it was the slowest role of the CTA by far, and
                    // with one oscillator whose amplitude ramps it still out-lasted the recurrence warp.
#pragma unroll
                    for (int i = 0; i < NOSC; ++i) {
                        const bool open = (open_mask >> i) & 1;
                        bool fast = false;
                        if (in_seg && open && osc[i].mode == OSC_MIP && osc[i].run == RUN_TABLE && !osc[i].p.timer &&
                            !osc[i].p_ramping && osc[i].dphase) {
                            const WaveDesc &w = c.waves[osc[i].wave];
                            const unsigned sz = w.size[osc[i].mm];
                            if (w.flags & kLooped) {
                                osc[i].ph = wrap_mod(osc[i].ph, (unsigned long long)sz << 24);
                                fast = true;
                            } else
                                fast = (osc[i].ph >> 24) <= (unsigned long long)(sz + kWavePre);
                            // the amplitude may ramp (a script that fades a voice does so across many
                            // fragments): its ramper is the only per-segment state left, wtosc.c:258
                            if (fast) {
                                ramp_prepare(osc[i].a, n);
                                osc[i].astep = osc[i].a.delta;
                            }
                        }
                        if (!fast) {
                            if (open) osc[i].finish();
                            if (in_seg) osc[i].prepare(c, n);
                            open_mask = in_seg ? (open_mask | (1u << i)) : (open_mask & ~(1u << i));
                        }
                    }
                    if (in_seg) {
                        if (FILT) filt.prepare(c, n);
                        pm.prepare(c, n);
                    }
                    // ---- publish the segment, advance in closed form ----
                    if (seg < kSplitSegs) {
                        if (in_seg) flags |= 1 << seg;
                        const int sb = (slot * kSplitSegs + seg);
#pragma unroll
                        for (int i = 0; i < NOSC; ++i) {
                            int *o = sm + L::oscp + ((sb * NOSC + i) * kOscWords) * 32 + lane;
                            // mode 0: silent; 1: coefficient table; 2: raw int16 taps; 3: raw taps, wrapped per sample
                            int mode = 0;
                            if (in_seg && osc[i].plain())
                                mode = osc[i].run == RUN_CHECK_LOOP ? 3 : (osc[i].cf ? 1 : 2);
                            o[0] = mode == 1 ? (int)(osc[i].cf - c.cpool) : mode ? (int)(osc[i].d - c.pool) : 0;
                            o[32] = (int)(unsigned)osc[i].ph;
                            o[64] = (int)(unsigned)(osc[i].ph >> 32);
                            o[96] = (int)osc[i].dph;
                            o[128] = osc[i].a.value;
                            o[160] = osc[i].astep;
                            o[192] = mode;
                            o[224] = (int)osc[i].wsize;
                            if (mode == 3) {
                                // wtosc.c:301-358, wrapped loop: sample k reads at (ph + k dph) mod M and the
                                // accumulator is left one increment past the last (reduced) read position
                                const unsigned long long M = (unsigned long long)osc[i].wsize << 24;
                                osc[i].ph = (osc[i].ph + (unsigned long long)osc[i].dph * (unsigned)(n - 1)) % M + osc[i].dph;
                                osc[i].a.value = wadd(osc[i].a.value, wmul(osc[i].astep, n));
                            } else if (mode) {          // wtosc.c:231-232 over n frames
                                osc[i].ph += (unsigned long long)osc[i].dph * (unsigned)n;
                                osc[i].a.value = wadd(osc[i].a.value, wmul(osc[i].astep, n));
                            }
                        }
                        if (FILT) {
                            int *o = sm + L::fp + (sb * 7) * 32 + lane;
                            o[0] = filt.f0; o[32] = filt.df; o[64] = filt.q.value; o[96] = filt.qstep;
                            o[128] = filt.lp; o[160] = filt.bp; o[192] = filt.hp;
                            if (in_seg) filt.q.value = wadd(filt.q.value, wmul(filt.qstep, n));
                        }
                        {
                            int *o = sm + L::pmp + (sb * 8) * 32 + lane;
                            o[0] = pm.vol.value; o[32] = pm.vstep; o[64] = pm.pan.value; o[96] = pm.pstep;
                            o[128] = pm.clamp ? 1 : 0;
                            // gains at rest (the usual case): the two channel gains of panmix.c:78-115
                            // are the same for every frame of the segment - computed here once per
                            // voice (lane = voice) instead of per frame and voice in stage C
                            const bool rest = pm.vstep == 0 && pm.pstep == 0;
                            const int vp = mulshr(pm.pan.value, pm.vol.value, 24);
                            int g0 = wsub(pm.vol.value, vp), g1 = wadd(pm.vol.value, vp);
                            if (pm.clamp) {
                                const int lim = (int)((unsigned)pm.vol.value << 1);
                                if (g0 > lim) g0 = lim;
                                if (g1 > lim) g1 = lim;
                            }
                            o[160] = g0; o[192] = g1; o[224] = rest ? 1 : 0;
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
            // stage C reads one word per voice: split | flags << 8 | (mixes into the CTA's home bus) << 16
            sm[L::flags + slot * 32 + lane] = flags | (split << 8) | ((valid && mybus == home) ? 1 << 16 : 0);
            if (lane == 0) { sm[L::meta + slot * 2] = f0; sm[L::meta + slot * 2 + 1] = fe - f0; }
            __syncwarp();
            if (lane == 0) mbar_arrive(&barP[slot]);
            trace(0, it, 1);
            f0 = fe;
            if (P.prof) busy += clock64() - tb;
        }
        if (valid) {
#pragma unroll
            for (int i = 0; i < NOSC; ++i)
                if ((open_mask >> i) & 1) osc[i].finish();
            if (in_seg) {
                if (FILT) filt.finish();
                pm.finish();
            }
            sp.st(0, alive);
#pragma unroll
            for (int i = 0; i < NOSC; ++i) osc[i].store(sp, 1 + 14 * i);
            if (FILT) {
                // d1 / d2 belong to the serial warp: store everything but them
                sp.st_ramp(L::filt_w, filt.cutoff); sp.st_ramp(L::filt_w + 4, filt.q);
                sp.st(L::filt_w + 8, filt.lp); sp.st(L::filt_w + 9, filt.bp); sp.st(L::filt_w + 10, filt.hp);
                sp.st(L::filt_w + 11, filt.f1);
            }
            pm.store(sp, L::pm_w);
        }
        // a staged copy still in flight may not outlive the CTA
        if (stage_n && set == 0 && lane == 0) mbar_wait(&s_mbar, 0);
    }

    // =====================================================================================
    // serial: filter12 recurrence, in place on the tile
    // =====================================================================================
    if (is_ser) {
        for (int it = 0; it < nfrag; ++it) {
            const int slot = it % R, par = (it / R) & 1;
            const long long tw = P.prof ? clock64() : 0;
            // A implies P: every helper acquired P before it wrote its part of the tile and released A
            mbar_wait(&barA[slot], par);
            const long long tb = P.prof ? clock64() : 0;
            trace(1, it, 0);
            const int n = sm[L::meta + slot * 2 + 1];
            const int split = sm[L::split + slot * 32 + lane];
            const int flags = sm[L::flags + slot * 32 + lane] & 0xff;
            int *t = sm + L::tile + slot * kMaxFrag * kTileStride + lane;
            for (int seg = 0; seg < kSplitSegs; ++seg) {
                const int a = seg ? split : 0, b = seg ? n : min(split, n);
                const bool live = a < b && ((flags >> seg) & 1);     // inactive: state frozen, output unused
                const int *o = sm + L::fp + ((slot * kSplitSegs + seg) * 7) * 32 + lane;
                const int df = o[32], qstep = o[96];
                const bool ramp = live && (df != 0 || qstep != 0);
                // warp-uniform choice: the constant-coefficient loop is the common case
                if (__any_sync(0xffffffffu, ramp)) {
                    if (live) filter_run<true>(t, a, b, d1, d2, o[0], df, o[64], qstep, o[128], o[160], o[192]);
                } else if (live)
                    filter_run<false>(t, a, b, d1, d2, o[0], 0, o[64], 0, o[128], o[160], o[192]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&barB[slot]);
            trace(1, it, 1);
            if (P.prof) { busy += clock64() - tb; busy2 += tb - tw; }
        }
        if (valid) {       // the recurrence state lives here
            sp.st(L::filt_w + 12, d1);
            sp.st(L::filt_w + 13, d2);
        }
    }

    // =====================================================================================
    // helpers: stage A (oscillators) of fragment it, stage C (panmix + bus) of fragment it - 1
    // =====================================================================================
    if (is_helper) {
        for (int it = 0; it <= nfrag; ++it) {
            // ---------------- stage A(it): lane = voice, frames sliced ----------------
            if (it < nfrag) {
                const int slot = it % R, par = (it / R) & 1;
                mbar_wait(&barP[slot], par);
                const long long tb = P.prof ? clock64() : 0;
                if (hq == 0) trace(2, it, 0);
                if (hq == NH - 1) trace(4, it, 0);
                const int n = sm[L::meta + slot * 2 + 1];
                if constexpr (LF) {
                    // lane = frame: unit u = (voice u % 32, frames [32 (u / 32), +32)) of this fragment
                    if (!tab_ready) { mbar_wait(&s_mbar, 0); tab_ready = true; }
                    int *tslot = sm + L::tile + slot * kMaxFrag * kTileStride;
                    const int nunits = ((n + 31) >> 5) << 5;
                    // B units at a time: all their taps are requested before the first is used
                    constexpr int B = RAW ? A2CU_LF_BATCH : 1;
#pragma unroll 1
                    for (int u0 = hq; u0 < nunits; u0 += B * NH) {
                        int vo[B], fr[B], kk[B], sg[B], acc[B];
                        bool on[B];
#pragma unroll
                        for (int j = 0; j < B; ++j) {
                            const int u = u0 + j * NH;
                            vo[j] = u & 31;
                            fr[j] = u < nunits ? (u & ~31) + lane : n;
                            const int split = sm[L::split + slot * 32 + vo[j]];
                            const int flags = sm[L::flags + slot * 32 + vo[j]];
                            sg[j] = fr[j] >= split ? 1 : 0;
                            kk[j] = fr[j] - (sg[j] ? split : 0);
                            on[j] = fr[j] < n && ((flags >> sg[j]) & 1);
                            acc[j] = 0;
                        }
#pragma unroll 1
                        for (int i = 0; i < NOSC; ++i) {
                            int mode[B], av[B];
                            unsigned p16[B], half[B];
                            const int *o[B];
                            RawTaps t0[B], t1[B];
#pragma unroll
                            for (int j = 0; j < B; ++j) {
                                mode[j] = 0;
                                if (!on[j]) continue;
                                o[j] = sm + L::oscp + (((slot * kSplitSegs + sg[j]) * NOSC + i) * kOscWords) * 32 + vo[j];
                                mode[j] = o[j][192];
                                if (!mode[j]) continue;         // silent segment of this oscillator
                                const unsigned dph = (unsigned)o[j][96];
                                unsigned long long ph = ((unsigned long long)(unsigned)o[j][64] << 32) | (unsigned)o[j][32];
                                ph += (unsigned long long)dph * (unsigned)kk[j];     // wtosc.c:231, k frames on
                                av[j] = wadd(o[j][128], wmul(o[j][160], kk[j]));
                                half[j] = dph >> 17;
                                if (RAW && mode[j] >= 2) {
                                    // wtosc.c:301-358: the wrapped loop reads sample k at (ph + k dph) mod M
                                    if (mode[j] == 3) ph = mod_wave(ph, (unsigned)o[j][224]);
                                    const int16_t *d = c.pool + o[j][0];
                                    p16[j] = (unsigned)(ph >> 16);
                                    t0[j] = hermite_fetch(d, p16[j]);
                                    t1[j] = hermite_fetch(d, p16[j] + half[j]);
                                } else
                                    p16[j] = (unsigned)(ph >> 16);
                            }
#pragma unroll
                            for (int j = 0; j < B; ++j) {
                                if (!mode[j]) continue;
                                int hv = 0;
                                if (mode[j] == 1) {
                                    const int cfo = o[j][0];
                                    const int srel = cfo - P.stage_begin;
                                    const int4 *cf = (srel >= 0 && srel < stage_n - 64) ? s_tab + srel : c.cpool + cfo;
                                    hv = hermite_cf_smem(cf, p16[j]) + hermite_cf_smem(cf, p16[j] + half[j]);
                                } else if (RAW) {
                                    const int16_t *d = c.pool + o[j][0];
                                    hv = hermite_eval(t0[j], d, p16[j]) + hermite_eval(t1[j], d, p16[j] + half[j]);
                                }
                                acc[j] = wadd(acc[j], mulshr(hv, av[j], 17));
                            }
                        }
#pragma unroll
                        for (int j = 0; j < B; ++j)
                            if (fr[j] < n) tslot[fr[j] * kTileStride + vo[j]] = acc[j];
                    }
                } else {
                    const int split = sm[L::split + slot * 32 + lane];
                    const int flags = sm[L::flags + slot * 32 + lane] & 0xff;
                    int *ta = sm + L::tile + slot * kMaxFrag * kTileStride + lane;
                    const int s0 = hq * kSlice, s1 = min(n, s0 + kSlice);
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
                                const int *o = sm + L::oscp + (((slot * kSplitSegs + seg) * NOSC + i) * kOscWords) * 32 + lane;
                                const int mode = o[192];
                                if (!mode) continue;            // silent segment of this oscillator
                                const unsigned dph = (unsigned)o[96];
                                unsigned long long ph = ((unsigned long long)(unsigned)o[64] << 32) | (unsigned)o[32];
                                const int astep = o[160];
                                const int av0 = wadd(o[128], wmul(astep, a - sa));
                                const unsigned half = dph >> 17;
                                if (mode == 1) {
                                    if (!tab_ready) { mbar_wait(&s_mbar, 0); tab_ready = true; }
                                    const int cfo = o[0];
                                    const int srel = cfo - P.stage_begin;
                                    // generic pointer: the staged copy in shared memory or the pool in global memory
                                    const int4 *cf = (srel >= 0 && srel < stage_n - 64) ? s_tab + srel : c.cpool + cfo;
                                    ph += (unsigned long long)dph * (unsigned)(a - sa);
                                    // all kSlice frames are evaluated (independent chains the scheduler can
                                    // overlap); frames past the segment end read table slack and are dropped
                                    // at the store (wtosc.c:226-233)
    #pragma unroll
                                    for (int k = 0; k < kSlice; ++k) {
                                        const unsigned p16 = (unsigned)((ph + (unsigned long long)dph * (unsigned)k) >> 16);
                                        const int hv = hermite_cf_smem(cf, p16) + hermite_cf_smem(cf, p16 + half);
                                        acc[k] = wadd(acc[k], mulshr(hv, wadd(av0, wmul(astep, k)), 17));
                                    }
                                } else if (RAW) {
                                    // Raw int16 taps from the pool (sampled waves too large for a table): the
                                    // gather that goes to HBM. The phase is closed-form, so the taps of all
                                    // frames of the slice are requested before the first is used.
                                    const int16_t *d = c.pool + o[0];
                                    unsigned p16s[kSlice];
                                    if (mode == 3) {
                                        const unsigned long long M = (unsigned long long)(unsigned)o[224] << 24;
                                        unsigned long long x = (ph + (unsigned long long)dph * (unsigned)(a - sa)) % M;
    #pragma unroll
                                        for (int k = 0; k < kSlice; ++k) {
                                            p16s[k] = (unsigned)(x >> 16);
                                            x = wrap_mod(x + dph, M);
                                        }
                                    } else {
                                        ph += (unsigned long long)dph * (unsigned)(a - sa);
    #pragma unroll
                                        for (int k = 0; k < kSlice; ++k)
                                            p16s[k] = (unsigned)((ph + (unsigned long long)dph * (unsigned)k) >> 16);
                                    }
    #pragma unroll
                                    for (int k = 0; k < kSlice; ++k) {
                                        // frames past the segment end are not fetched (no table slack in the pool)
                                        const bool in_range = a + k < b;
                                        const int hv = in_range ? hermite(d, p16s[k]) + hermite(d, p16s[k] + half) : 0;
                                        acc[k] = wadd(acc[k], mulshr(hv, wadd(av0, wmul(astep, k)), 17));
                                    }
                                }
                            }
                        }
    #pragma unroll
                        for (int k = 0; k < kSlice; ++k)
                            if (a + k < b) ta[(a + k) * kTileStride] = acc[k];
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&barA[slot]);
                if (hq == 0) trace(2, it, 1);
                if (hq == NH - 1) trace(4, it, 1);
                if (P.prof) busy += clock64() - tb;
            }
            // ---------------- stage C(it - 1): panmix + bus sum, lane = frame ----------------
            if (it >= 1) {
                const int jt = it - 1;
                const int slot = jt % R, par = (jt / R) & 1;
                mbar_wait(FILT ? &barB[slot] : &barA[slot], par);
                const long long tb = P.prof ? clock64() : 0;
                if (hq == 0) trace(3, jt, 0);
                if (hq == NH - 1) trace(5, jt, 0);
                {
                    // lane = voice, my slice of the fragment's frames (the same slice as stage A).
                    // The voice's gains live in registers, a frame costs one conflict-free tile read,
                    // two 64-bit multiplies and two warp reductions (redux.sync: PROCADD "+=" over
                    // the 32 voices, panmix.c:104-105), and the slice's sums are added to the set's
                    // window accumulator by their one owner - no atomics, no per-voice parameter
                    // fetches per frame.
                    const int f0 = sm[L::meta + slot * 2];
                    const int n = sm[L::meta + slot * 2 + 1];
                    const int info = sm[L::flags + slot * 32 + lane];
                    const int split = (info >> 8) & 0xff;
                    const bool athome = (info >> 16) != 0;
                    const int *pm0 = sm + L::pmp + ((slot * kSplitSegs + 0) * 8) * 32 + lane;
                    const int *pm1 = sm + L::pmp + ((slot * kSplitSegs + 1) * 8) * 32 + lane;
                    const int ga0 = pm0[160], gb0 = pm0[192], ga1 = pm1[160], gb1 = pm1[192];
                    const bool rest0 = pm0[224] != 0, rest1 = pm1[224] != 0;
                    const int *tin = sm + L::tile + slot * kMaxFrag * kTileStride + lane;
                    const int s0 = hq * kSlice, s1 = min(n, s0 + kSlice);
                    int keep0 = 0, keep1 = 0;
#pragma unroll
                    for (int k = 0; k < kSlice; ++k) {
                        const int f = s0 + k;
                        if (f >= s1) break;                 // warp-uniform
                        const int seg = f >= split ? 1 : 0;
                        const bool active = ((info >> seg) & 1) != 0;
                        int v0 = seg ? ga1 : ga0, v1 = seg ? gb1 : gb0;
                        if (active && !(seg ? rest1 : rest0)) {     // ramping gains: panmix.c:78-115 per frame
                            const int *o = seg ? pm1 : pm0;
                            const int kk = f - (seg ? split : 0);
                            const int vol = wadd(o[0], wmul(o[32], kk));
                            const int pan = wadd(o[64], wmul(o[96], kk));
                            const int vp = mulshr(pan, vol, 24);
                            v0 = wsub(vol, vp); v1 = wadd(vol, vp);
                            if (o[128]) {
                                const int lim = (int)((unsigned)vol << 1);
                                if (v0 > lim) v0 = lim;
                                if (v1 > lim) v1 = lim;
                            }
                        }
                        const int in = tin[f * kTileStride];
                        int o0 = active ? mulshr(in, v0, 24) : 0;
                        int o1 = active ? mulshr(in, v1, 24) : 0;
                        if (active && !athome) {            // a voice of another bus (group): straight to it
                            int *gp = P.acc + ((size_t)sm[L::bus + lane] * W + f0 + f) * 2;
                            atomicAdd(gp, o0);
                            atomicAdd(gp + 1, o1);
                            o0 = o1 = 0;
                        }
                        const int r0 = __reduce_add_sync(0xffffffffu, o0);
                        const int r1 = __reduce_add_sync(0xffffffffu, o1);
                        if (lane == k) { keep0 = r0; keep1 = r1; }
                    }
                    if (lane < s1 - s0) {
                        int2 *sa = reinterpret_cast<int2 *>(wacc + (f0 + s0 + lane) * 2);
                        int2 cur = *sa;
                        cur.x = wadd(cur.x, keep0); cur.y = wadd(cur.y, keep1);
                        *sa = cur;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&barC[slot]);
                if (hq == 0) trace(3, jt, 1);
                if (hq == NH - 1) trace(5, jt, 1);
                if (P.prof) busy2 += clock64() - tb;
            }
        }
    }

    if (P.prof && lane == 0) {
        // [0] control [1] serial compute [2] stage A [3] stage C [4] serial wait, [5] fragments, summed over CTAs
        if (is_ctl) atomicAdd(P.prof + 0, (unsigned long long)busy);
        else if (is_ser) { atomicAdd(P.prof + 1, (unsigned long long)busy); atomicAdd(P.prof + 4, (unsigned long long)busy2); }
        else if (is_helper && hq == 1) {
            atomicAdd(P.prof + 2, (unsigned long long)busy);
            atomicAdd(P.prof + 3, (unsigned long long)busy2);
            if (set == 0) atomicAdd(P.prof + 5, (unsigned long long)nfrag);
        }
    }
    __syncthreads();                        // all stage C work of this CTA is done
    gmark(1, false);
    for (int s2 = 0; s2 < VS; ++s2) {
        const int first = (blockIdx.x * VS + s2) * 32;
        if (first >= P.nvoices) break;
        int *ga = P.acc + (size_t)P.bus_of[first] * W * 2;
        const int *wa = wacc_all + s2 * kSplitMaxWin * 2;
        for (int i = tid; i < W * 2; i += WR::threads) {
            const int val = wa[i];
            if (val) atomicAdd(ga + i, val);
        }
    }
    __syncthreads();                        // every bus reduction of this CTA has been issued
    gmark(2, false);
    if (P.prof && tid == 0) {               // [6] prologue, [7] pipeline + state store, per CTA (a2cu_split_profile)
        atomicAdd(P.prof + 6, (unsigned long long)(t_loop - t_entry));
        atomicAdd(P.prof + 7, (unsigned long long)(clock64() - t_loop));
    }
    // ---- fused root stage: the last CTA to get here owns the finished root bus ----
    if (P.fuse_root) {
        __shared__ int s_last;
        if (tid == 0) {
            __threadfence();                // ... and is visible before the ticket is taken
            s_last = atomicAdd(P.fuse_counter, 1u) == gridDim.x - 1 ? 1 : 0;
        }
        __syncthreads();
        if (s_last) {
            // (the root stage reads the bus with ld.cg, i.e. from L2 where the other CTAs' reductions
            // were performed; thread 0's fence + ticket + the barrier above order them before us)
            MixParams M;
            M.acc = P.acc; M.W = P.W; M.buffer = P.buffer; M.ngroups = 0; M.channels = P.fuse_channels;
            M.nsplits = P.nsplits;          // root wake-ups cut the root panmix's segments too
            for (int i = 0; i < kMaxSplits; ++i) M.splits[i] = P.splits[i];
            M.gstate = nullptr; M.rstate = P.fuse_rstate; M.ev = nullptr; M.nev = 0;
            M.master = P.fuse_master; M.root_stage = P.fuse_root_stage; M.clear = 1; M.general = 0;
            M.out_fmt = P.fuse_out_fmt;
            // tail timings for a2cu_split_trace: role 5, fragments 60..63 = start, after the previous
            // window's root stage, after publish / root stage, (unused)
            auto tmark = [&](int k) {
                if (P.prof && tid == 0) P.prof[8 + ((5 * 64 + 60 + k) * 2)] = (unsigned long long)(clock64() - t_loop);
            };
            tmark(0);
            if (P.xchg.world > 1 && P.xchg.lag) {
                // sharded + pipelined: finish the PREVIOUS window (its rows arrived long ago), then
                // publish this one without waiting for anybody (read-then-publish keeps two buffer
                // halves enough: a peer publishes window k + 1 only after it saw our window k)
                if (P.xchg.prev_valid)
                    xchg_finish_previous(P.xchg, P.fuse_rstate, P.fuse_channels, P.fuse_root_stage, tid, WR::threads,
                                         P.fuse_out_fmt);
                tmark(1);
                xchg_publish(P.xchg, P.acc, P.W, tid, WR::threads, true);
                tmark(2);
            } else {
                // sharded render: the root bus of all ranks is summed here, through NVLink peer memory
                if (P.xchg.world > 1) xchg_root_bus(P.xchg, P.acc, P.W, tid, WR::threads);
                root_stage(M, tid, WR::threads, true);
                tmark(2);
            }
            gmark(3, false);
            if (tid == 0) *P.fuse_counter = 0u;     // ready for the next launch
        }
    }
}

}  // namespace a2cu
