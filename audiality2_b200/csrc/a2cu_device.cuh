// a2cu_device.cuh - device-side DSP primitives and voice-unit templates.
//
// One CUDA thread owns one voice. A voice structure ("chain") is a compile-time
// list of unit templates; the scratch channels that the reference keeps in
// st->scratch[nestlevel] (core.c:364-395) live in two registers. Per-voice
// state is structure-of-arrays in HBM: word w of slot s is state[w * stride + s],
// so a warp's load of one word is one coalesced 128-byte transaction.
//
// All audio arithmetic is integer and must match the reference bit for bit;
// signed overflow is performed on unsigned operands (wraps like gcc/x86-64).
// Reference file:line is cited at each function.
#pragma once
#include <stdint.h>

namespace a2cu {

#define A2CU_DEV __device__ __forceinline__

constexpr int kMaxFrag = 64;      // A2_MAXFRAG, audiality2.h.cmake:50
constexpr int kMipLevels = 10;    // a2_waves.h:33
constexpr int kWavePre = 1;       // a2_waves.h:61
constexpr int kMaxPhInc = 512;    // a2_waves.h:58

// ---- wrapping helpers ------------------------------------------------------
A2CU_DEV int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
A2CU_DEV int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
A2CU_DEV int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
A2CU_DEV int mulshr(int a, int b, int sh) {       // (int64)a * b >> sh, truncated
    return (int)(((long long)a * (long long)b) >> sh);
}

// ---- wave descriptors (read side of A2_wave, a2_waves.h:88-103) -------------
enum WaveType { W_OFF = 0, W_NOISE = 1, W_WAVE = 2, W_MIPWAVE = 3 };
constexpr unsigned kLooped = 0x100;               // A2_LOOPED, a2_waves.h:108

struct WaveDesc {
    int type;
    unsigned flags;
    unsigned period;
    unsigned size[kMipLevels];      // excluding pads
    unsigned offset[kMipLevels];    // index of data[level][A2_WAVEPRE] in the pool
    int coff[kMipLevels];           // index of sample 0's entry in the Hermite coefficient pool, -1: none
};

// Everything a unit needs besides its own state.
struct Ctx {
    const WaveDesc *waves;
    const int16_t *pool;            // all wave data, int16, pads included
    const int4 *cpool;              // two-stage Hermite coefficients {d0, a, b, c} per sample
    const unsigned *ptab;           // 64 x {base, coeff}, pitch.c:70-96
    const int *fmsine;              // 2048 x {sine[i] | (sine[i + 1] - sine[i]) << 16} (fm.c:486-501, packed by the host)
    const int *f12tab;              // filter12 coefficient of every 21-bit (shift, fraction) pitch code
    int samplerate;
};

// Exact C-style (truncating) num / den for |num| < 2^52, den > 0: a double
// estimate corrected with the integer remainder. Replaces the ~100-instruction
// software s64 division in the ramper prologue (a2_dsp.h:137).
A2CU_DEV long long div_trunc(long long num, int den) {
    long long q = (long long)((double)num / (double)den);
    long long r = num - q * den;
    if (num >= 0) {
        if (r < 0) --q; else if (r >= den) ++q;
    } else {
        if (r > 0) ++q; else if (r <= -(long long)den) --q;
    }
    return q;
}
// x % m, cheap when x < 2m (the usual case for a looped phase accumulator)
A2CU_DEV unsigned long long wrap_mod(unsigned long long x, unsigned long long m) {
    if (x < m) return x;
    x -= m;
    if (x < m) return x;
    return x % m;
}

// ---- a2_dsp.h ---------------------------------------------------------------
struct Ramp { int value, target, delta, timer; };

// a2_dsp.h:128-149
A2CU_DEV void ramp_prepare(Ramp &r, int frames) {
    if (!r.timer) {
        r.value = r.target;
        r.delta = 0;
    } else if (frames <= (r.timer >> 8)) {
        r.delta = (int)div_trunc((long long)wsub(r.target, r.value) << 8, r.timer);
        r.timer -= frames << 8;
    } else {
        r.delta = wsub(r.target, r.value) / frames;
        r.timer = 0;
    }
}
// a2_dsp.h:152-155
A2CU_DEV void ramp_run(Ramp &r, int frames) { r.value = wadd(r.value, wmul(r.delta, frames)); }
// a2_dsp.h:161-170
A2CU_DEV void ramp_set(Ramp &r, int target, int start, int dur) {
    r.target = (int)((unsigned)target << 8);
    r.timer = dur + start;
    if (r.timer < 256)
        r.value = r.target;
    else
        r.value = wadd(r.value, wmul(r.delta, start) >> 8);
}
// a2_dsp.h:121-125
A2CU_DEV void ramp_init(Ramp &r, int v) {
    r.value = r.target = (int)((unsigned)v << 8);
    r.delta = r.timer = 0;
}

// a2_dsp.h:37-42
A2CU_DEV int noise_next(unsigned &st) {
    st = st * 1566083941u + 1u;
    return (int)((st * (st >> 16)) >> 16);
}

// a2_dsp.h:50-55
A2CU_DEV int lerp16(const int16_t *d, unsigned ph) {
    int i = ph >> 8;
    int x = ph & 0xff;
    return (d[i] * (256 - x) + d[i + 1] * x) >> 8;
}

// The same interpolation from a table of {d[i], d[i + 1] - d[i]} pairs: one 4-byte load instead of two
// 2-byte ones, one multiply instead of two. Exact: (256 d0 + (d1 - d0) x) >> 8 = d0 + (((d1 - d0) x) >> 8)
// for an arithmetic shift, and the difference of two neighbours of the FM sine fits 16 bits.
A2CU_DEV int lerp16_packed(const int *t, unsigned ph) {
    const int w = t[ph >> 8];
    const int x = (int)(ph & 0xff);
    return (int)(short)w + (((w >> 16) * x) >> 8);
}

// a2_dsp.h:64-74 on raw taps
A2CU_DEV int hermite4(int dm, int d0, int d1, int d2, unsigned ph) {
    int x = (int)((ph & 0xff) << 7);
    int c = (d1 - dm) >> 1;
    int a = (3 * (d0 - d1) + d2 - dm) >> 1;
    int b = dm - d0 + c - a;
    a = wmul(a, x) >> 15;
    a = wmul(a + b, x) >> 15;
    return d0 + (wmul(a + c, x) >> 15);
}
// The four taps d[i-1..i+2] are 8 consecutive bytes at a 2-byte aligned address. For large sampled
// waves (no coefficient table) every lane gathers somewhere else in HBM and the kernel is bound by
// sector requests through L1, so the taps are fetched with as few requests as alignment allows: the
// aligned 16-byte chunk that holds the first tap, plus - only for the 3 of 8 alignments where the
// taps run past it - the next 8 bytes (1.4 requests per tap on average; round 1: four 2-byte loads,
// then two 8-byte loads). The pool pads each level with A2_WAVEPRE / A2_WAVEPOST samples and the
// arena has slack behind the last wave, so these reads stay inside the allocation.
// Split in two so that a caller can request the taps of several samples before it uses the first
// (the closed-form phase makes every address known up front).
struct RawTaps { uint4 c; unsigned long long nx; };
A2CU_DEV RawTaps hermite_fetch(const int16_t *d, unsigned ph) {
    const unsigned long long a = (unsigned long long)(d + (int)(ph >> 8) - 1);
    RawTaps t;
    t.c = __ldg(reinterpret_cast<const uint4 *>(a & ~15ull));
    t.nx = (a & 15ull) > 8 ? __ldg(reinterpret_cast<const unsigned long long *>((a & ~15ull) + 16)) : 0ull;
    return t;
}
// The taps are halfwords e .. e + 3 of the 24-byte window {c, nx}, e = byte offset / 2: two rounds of
// word selects and two funnel shifts, branch-free (the first version branched on the alignment and
// used 64-bit shifts: ~27 instructions per tap and a divergent warp; this is 12).
A2CU_DEV int hermite_eval(const RawTaps &t, const int16_t *d, unsigned ph) {
    const unsigned o = (unsigned)((unsigned long long)(d + (int)(ph >> 8) - 1) & 15ull);     // 0, 2, ... 14
    const unsigned n0 = (unsigned)t.nx, n1 = (unsigned)(t.nx >> 32);
    const bool up2 = (o & 8u) != 0, up1 = (o & 4u) != 0;
    const unsigned a0 = up2 ? t.c.z : t.c.x, a1 = up2 ? t.c.w : t.c.y, a2 = up2 ? n0 : t.c.z, a3 = up2 ? n1 : t.c.w;
    const unsigned b0 = up1 ? a1 : a0, b1 = up1 ? a2 : a1, b2 = up1 ? a3 : a2;
    const unsigned sh = (o & 2u) << 3;                  // 0 or 16
    const unsigned lo = __funnelshift_r(b0, b1, sh), hi = __funnelshift_r(b1, b2, sh);
    return hermite4((short)lo, (int)lo >> 16, (short)hi, (int)hi >> 16, ph);
}
A2CU_DEV int hermite(const int16_t *d, unsigned ph) { return hermite_eval(hermite_fetch(d, ph), d, ph); }
// a2_Hermite2 (a2_dsp.h:91-98) on precomputed a2_Hermite2c coefficients
// (a2_dsp.h:83-89): term for term the same integers as a2_Hermite, but one
// 16-byte load instead of four unaligned int16 loads.
A2CU_DEV int hermite_cf(const int4 *cf, unsigned ph) {
    const int4 e = __ldg(cf + (int)(ph >> 8));      // {d0, a, b, c}
    int x = (int)((ph & 0xff) << 7);
    int a = wmul(e.y, x) >> 15;
    a = wmul(a + e.z, x) >> 15;
    return e.x + (wmul(a + e.w, x) >> 15);
}

// pitch.c:57-67; the shift count is taken & 31 like the x86-64 build does
A2CU_DEV unsigned p2i(const unsigned *ptab, int pitch) {
    int n = pitch & 0xffff;
    int oct = pitch >> 16;
    unsigned base = __ldg(ptab + 2 * (n >> 10));
    unsigned coeff = __ldg(ptab + 2 * (n >> 10) + 1);
    unsigned dph = coeff * (unsigned)(n & 0x3ff);
    dph >>= 2;
    dph += base;
    return dph >> ((7 - oct) & 31);
}

// filter12.c:65-72, f12_pitch2coeff(): `float f = a2_P2I(cutoff >> 8) * (261.626f / 16777216.0f)`, then
// the double `sin` of the host libm. a2_P2I (pitch.c:57-67) only looks at the 16 fraction bits of
// the pitch and at (7 - octave) & 31, so for one sample rate the function has 32 x 65536 distinct
// arguments: the host evaluates all of them once with ITS libm (the reference's own expression,
// a2cu_engine.cu f12_table) and the device looks the result up. Bit-exact by construction: no
// device libm is involved.
A2CU_DEV int f12_coeff(const Ctx &c, int cutoff_value) {
    const int pitch = cutoff_value >> 8;
    return __ldg(c.f12tab + ((((7 - (pitch >> 16)) & 31) << 16) | (pitch & 0xffff)));
}

// Same interpolation through a generic pointer: the table staged in shared
// memory by TMA, or the global pool (L1-cached)
A2CU_DEV int hermite_cf_smem(const int4 *tab, unsigned ph) {
    const int4 e = tab[(int)(ph >> 8)];
    int x = (int)((ph & 0xff) << 7);
    int a = wmul(e.y, x) >> 15;
    a = wmul(a + e.z, x) >> 15;
    return e.x + (wmul(a + e.w, x) >> 15);
}

// ---- TMA (1-D bulk copy) + mbarrier, inline PTX --------------------------------
A2CU_DEV unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
A2CU_DEV void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
A2CU_DEV void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
A2CU_DEV void tma_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Wait for the phase with the given parity. try_wait suspends the warp in hardware until the phase
// completes or the time hint expires; the explicit (long) hint matters: with the short default a
// waiting warp keeps re-issuing the poll, and since the scheduler prefers higher warp ids, spinning
// warps starve lower-numbered working warps of their sub-partition (measured: a helper lagging
// 3.6 k cycles behind a barrier that was already complete, profiles/r02_split_timeline_spin.txt).
A2CU_DEV void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
            : "memory");
    }
}

// ---- state I/O --------------------------------------------------------------
struct StatePtr {
    int *base;          // word 0 of this voice (already offset by slot)
    size_t stride;      // slots per word row
    A2CU_DEV int ld(int w) const { return base[(size_t)w * stride]; }
    A2CU_DEV void st(int w, int v) const { base[(size_t)w * stride] = v; }
    A2CU_DEV void ld_ramp(int w, Ramp &r) const {
        r.value = ld(w); r.target = ld(w + 1); r.delta = ld(w + 2); r.timer = ld(w + 3);
    }
    A2CU_DEV void st_ramp(int w, const Ramp &r) const {
        st(w, r.value); st(w + 1, r.target); st(w + 2, r.delta); st(w + 3, r.timer);
    }
};

// =============================================================================
// wtosc (src/units/wtosc.c)
// =============================================================================
enum OscMode { OSC_OFF = 0, OSC_NOISE = 1, OSC_NOMIP = 2, OSC_MIP = 3 };
// per-segment inner loop variants
enum OscRun { RUN_SILENT = 0, RUN_TABLE = 1, RUN_NOISE = 2, RUN_CHECK_LOOP = 3, RUN_CHECK_END = 4 };

template <bool ADD, bool WIREOUT>
struct WtOsc {
    static constexpr int kWords = 14;
    static constexpr int kNIn = 0, kNOut = 1;
    static constexpr bool kUsesFm = false;
    // persistent state (wtosc.c:66-80)
    Ramp p, a;
    unsigned dphase;
    unsigned long long phase;
    int noise, p_ramping;
    int wave;           // index into Ctx::waves or -1
    int mode;           // OscMode
    // segment-local
    const int16_t *d;
    const int4 *cf;     // coefficient table of the current level or nullptr
    unsigned long long ph;
    unsigned dph;
    unsigned wsize;
    unsigned nstate;    // LCG state for this segment's noise draws (EV_SEED)
    int astep;
    int mm, run;

    A2CU_DEV void load(const StatePtr &s, int w) {
        s.ld_ramp(w, p); s.ld_ramp(w + 4, a);
        dphase = (unsigned)s.ld(w + 8);
        phase = (unsigned)s.ld(w + 9) | ((unsigned long long)(unsigned)s.ld(w + 10) << 32);
        noise = s.ld(w + 11);
        int x = s.ld(w + 12);
        p_ramping = s.ld(w + 13);
        wave = x >> 8; mode = x & 0xff;
        run = RUN_SILENT; astep = 0; mm = 0; d = nullptr; cf = nullptr; ph = 0; dph = 0; wsize = 0; nstate = 0;
    }
    A2CU_DEV void store(const StatePtr &s, int w) const {
        s.st_ramp(w, p); s.st_ramp(w + 4, a);
        s.st(w + 8, (int)dphase);
        s.st(w + 9, (int)(unsigned)phase); s.st(w + 10, (int)(unsigned)(phase >> 32));
        s.st(w + 11, noise);
        s.st(w + 12, (wave << 8) | mode);
        s.st(w + 13, p_ramping);
    }

    // wtosc.c:378-387
    A2CU_DEV void set_phase(const Ctx &c, int phv, unsigned sst) {
        if (wave < 0) { phase = 0; return; }
        phv = wadd(phv, (int)((sst * (dphase >> 8)) >> 8));
        phase = (unsigned long long)(((long long)phv * (long long)c.waves[wave].period) << 8);
    }
    // wtosc.c:390-423; arg = transpose + basepitch, sst = waketime & 0xff
    A2CU_DEV void init(const Ctx &c, int arg, unsigned sst) {
        noise = 0; wave = -1;
        ramp_init(a, 0);
        ramp_init(p, arg);
        dphase = p2i(c.ptab, p.value >> 8);
        p_ramping = 0;
        set_phase(c, 0, sst);
        mode = OSC_OFF;
    }
    // wtosc.c:433-504. Values are pre-cooked by the host: reg 0 carries the
    // device wave index (-1 = off/invalid), reg 1 the pitch incl. transpose
    // and basepitch.
    A2CU_DEV void write(const Ctx &c, int reg, int v, int start, int dur) {
        switch (reg) {
        case 0:
            wave = v;
            if (v < 0) mode = OSC_OFF;
            else {
                int t = c.waves[v].type;
                mode = t == W_NOISE ? OSC_NOISE : t == W_WAVE ? OSC_NOMIP :
                       t == W_MIPWAVE ? OSC_MIP : OSC_OFF;
                if (mode == OSC_OFF) wave = -1;
            }
            break;
        case 1:
            ramp_set(p, v, start, dur);
            if (!dur) p_ramping = 1;
            break;
        case 2: ramp_set(a, v, start, dur); break;
        case 3: set_phase(c, v, (unsigned)start); break;
        }
    }
    // wtosc.c:89-105
    A2CU_DEV void run_pitch(const Ctx &c, int frames) {
        ramp_prepare(p, frames);
        if (dphase && (!p.timer && !p_ramping)) return;
        unsigned lastv = (unsigned)p.value;
        ramp_run(p, frames);
        p_ramping = p.delta;
        dphase = p2i(c.ptab, (int)((lastv + (unsigned)p.value) >> 9));
    }
    // Segment prologue: everything the reference does per Process() call
    // before/around the sample loop (wtosc.c:108-126, 129-137, 239-286, 301-358)
    A2CU_DEV void prepare(const Ctx &c, int frames) {
        run = RUN_SILENT; astep = 0; mm = 0;
        if (mode == OSC_MIP || mode == OSC_NOMIP) {
            const WaveDesc &w = c.waves[wave];
            if (!w.size[0]) {           // unloaded while playing, wtosc.c:168-183
                wave = -1; mode = OSC_OFF;
                return;
            }
        }
        switch (mode) {
        case OSC_OFF:
            ramp_prepare(p, frames); ramp_prepare(a, frames);
            ramp_run(p, frames); ramp_run(a, frames);
            return;
        case OSC_NOISE:
            run_pitch(c, frames);
            ramp_prepare(a, frames);
            run = RUN_NOISE; astep = a.delta;
            return;
        case OSC_MIP: {
            const WaveDesc &w = c.waves[wave];
            run_pitch(c, frames);
            unsigned e = ((dphase + 255) >> 8) * w.period;
            ramp_prepare(a, frames);
            int m = 0;
            for (; (e > (unsigned)(kMaxPhInc << 8)) && (m < kMipLevels - 1); ++m) e >>= 1;
            mm = m;
            ph = phase >> m;
            dph = (unsigned)(((unsigned long long)dphase * w.period) >> m);
            if (w.flags & kLooped)
                ph = wrap_mod(ph, (unsigned long long)w.size[m] << 24);
            else if ((ph >> 24) > (unsigned long long)(w.size[m] + kWavePre))
                return;                 // all played: silence, nothing advances
            if (dph > (unsigned)(kMaxPhInc << 16)) {
                ph += (unsigned long long)dph * (unsigned)frames;
                phase = ph << m;
                ramp_run(a, frames);
                return;                 // out of range: muted
            }
            d = c.pool + w.offset[m];
            cf = w.coff[m] >= 0 ? c.cpool + w.coff[m] : nullptr;
            run = RUN_TABLE; astep = a.delta; wsize = 0;
            return;
        }
        case OSC_NOMIP: {
            const WaveDesc &w = c.waves[wave];
            run_pitch(c, frames);
            unsigned long long dp = (unsigned long long)dphase * w.period;
            ramp_prepare(a, frames);
            d = c.pool + w.offset[0];
            cf = w.coff[0] >= 0 ? c.cpool + w.coff[0] : nullptr;
            if (dp >> 32) {
                phase += dp * (unsigned)frames;
                ramp_run(a, frames);
                return;
            }
            dph = (unsigned)dp;
            if (dp > (unsigned long long)(kMaxPhInc << 16)) {
                ph = phase;
                wsize = w.size[0];
                run = (w.flags & kLooped) ? RUN_CHECK_LOOP : RUN_CHECK_END;
                astep = a.delta;
                return;
            }
            if (w.flags & kLooped) {
                unsigned m32 = w.size[0] << 24;     // 32-bit, as written (wtosc.c:346)
                if (m32) phase %= m32;              // (reference divides by zero here)
            } else if ((phase >> 24) > (unsigned long long)(w.size[0] + kWavePre))
                return;
            ph = phase;
            run = RUN_TABLE; astep = a.delta; wsize = 0;
            return;
        }
        }
    }
    // One output sample. wtosc.c:200-236 (table), :139-151 (noise)
    A2CU_DEV void sample(const Ctx &c, int &s0, int &s1, int &o0, int &o1) {
        int v = 0;
        if (run == RUN_TABLE || run >= RUN_CHECK_LOOP) {
            bool live = true;
            if (run == RUN_CHECK_LOOP)
                ph %= (unsigned long long)wsize << 24;
            else if (run == RUN_CHECK_END && (ph >> 24) >= wsize) {
                run = RUN_SILENT; astep = 0; live = false;
                phase = ph;
            }
            if (live) {
                unsigned p16 = (unsigned)(ph >> 16);
                unsigned dp16 = dph >> 16;
                int h = cf ? hermite_cf(cf, p16) + hermite_cf(cf, p16 + (dp16 >> 1))
                           : hermite(d, p16) + hermite(d, p16 + (dp16 >> 1));
                v = mulshr(h, a.value, 17);
                ph += dph;
                a.value = wadd(a.value, astep);
            }
        } else if (run == RUN_NOISE) {
            unsigned long long nph = phase + dphase;
            if ((dphase >= (1u << 23)) || ((nph ^ phase) >> 23))
                noise = noise_next(nstate) - 32767;
            phase = nph;
            v = wmul(noise, a.value >> 10) >> 6;
            a.value = wadd(a.value, astep);
        }
        if (WIREOUT) o0 = wadd(o0, v);
        else if (ADD) s0 = wadd(s0, v);
        else s0 = v;
    }
    // The shared noise LCG (a2_dsp.h:37-42, wtosc.c:135-144) is advanced in
    // tree-walk order on the host; each noise segment gets its start state.
    A2CU_DEV void seed(unsigned s) { nstate = s; }
    // True when the current segment is a wavetable loop without a state change in the middle: the
    // unchecked loop, or the per-sample wrapped loop of a looped non-mipmapped wave played faster
    // than A2_MAXPHINC (wtosc.c:301-358) - the large-sampled-wave case whose gather goes to HBM.
    A2CU_DEV bool plain() const { return run == RUN_TABLE || run == RUN_CHECK_LOOP; }
    // sample() specialised for plain(): no liveness tests, so the compiler can overlap the Hermite
    // gathers of consecutive frames (wtosc.c:226-233) - for sampled waves that means several HBM
    // sectors in flight per voice instead of one frame's worth.
    A2CU_DEV void sample_fast(const Ctx &, int &s0, int &s1, int &o0, int &o1) {
        if (run == RUN_CHECK_LOOP) ph = wrap_mod(ph, (unsigned long long)wsize << 24);
        unsigned p16 = (unsigned)(ph >> 16);
        int h = cf ? hermite_cf(cf, p16) + hermite_cf(cf, p16 + (dph >> 17))
                   : hermite(d, p16) + hermite(d, p16 + (dph >> 17));
        int v = mulshr(h, a.value, 17);
        ph += dph;
        a.value = wadd(a.value, astep);
        if (WIREOUT) o0 = wadd(o0, v);
        else if (ADD) s0 = wadd(s0, v);
        else s0 = v;
    }
    // Segment epilogue: write the phase accumulator back (wtosc.c:283-285)
    A2CU_DEV void finish() {
        if (run == RUN_TABLE || run >= RUN_CHECK_LOOP) phase = ph << mm;
        run = RUN_SILENT; astep = 0;
    }
};

// =============================================================================
// panmix (src/units/panmix.c)
// =============================================================================
template <int NIN, int NOUT, bool ADD, bool WIREOUT>
struct PanMix {
    static constexpr int kWords = 8;
    static constexpr bool kUsesFm = false;
    Ramp vol, pan;
    int vstep, pstep;
    bool clamp;

    A2CU_DEV void load(const StatePtr &s, int w) {
        s.ld_ramp(w, vol); s.ld_ramp(w + 4, pan);
        vstep = pstep = 0; clamp = false;
    }
    A2CU_DEV void store(const StatePtr &s, int w) const { s.st_ramp(w, vol); s.st_ramp(w + 4, pan); }
    A2CU_DEV void init(const Ctx &, int, unsigned) { ramp_init(vol, 65536); ramp_init(pan, 0); }
    A2CU_DEV void write(const Ctx &, int reg, int v, int start, int dur) {
        ramp_set(reg == 0 ? vol : pan, v, start, dur);
    }
    A2CU_DEV void prepare(const Ctx &, int frames) {
        if (NIN == 1 && NOUT == 1) {        // panmix.c:49-64: pan untouched
            ramp_prepare(vol, frames);
            vstep = vol.delta; pstep = 0;
            return;
        }
        // clamp variant is picked per call, before Prepare (panmix.c:117-135)
        clamp = pan.target > 0xffffff || pan.target < -0xffffff ||
                pan.value > 0xffffff || pan.value < -0xffffff;
        ramp_prepare(vol, frames);
        ramp_prepare(pan, frames);
        vstep = vol.delta; pstep = pan.delta;
    }
    A2CU_DEV void sample(const Ctx &, int &s0, int &s1, int &o0, int &o1) {
        int r0, r1 = 0;
        if (NIN == 1 && NOUT == 1) {
            r0 = mulshr(s0, vol.value, 24);
        } else {
            int v = vol.value;
            int vp = mulshr(pan.value, v, 24);
            int v0 = wsub(v, vp), v1 = wadd(v, vp);
            if (clamp) {
                int lim = (int)((unsigned)v << 1);
                if (v0 > lim) v0 = lim;
                if (v1 > lim) v1 = lim;
            }
            if (NIN == 1) {
                r0 = mulshr(s0, v0, 24); r1 = mulshr(s0, v1, 24);
            } else if (NOUT == 1) {
                r0 = (int)(((long long)s0 * v0 + (long long)s1 * v1) >> 25);
            } else {
                r0 = mulshr(s0, v0, 24); r1 = mulshr(s1, v1, 24);
            }
            pan.value = wadd(pan.value, pstep);
        }
        vol.value = wadd(vol.value, vstep);
        if (WIREOUT) { o0 = wadd(o0, r0); if (NOUT == 2) o1 = wadd(o1, r1); }
        else if (ADD) { s0 = wadd(s0, r0); if (NOUT == 2) s1 = wadd(s1, r1); }
        else { s0 = r0; if (NOUT == 2) s1 = r1; }
    }
    A2CU_DEV bool plain() const { return true; }
    A2CU_DEV void sample_fast(const Ctx &c, int &s0, int &s1, int &o0, int &o1) { sample(c, s0, s1, o0, o1); }
    A2CU_DEV void seed(unsigned) {}
    A2CU_DEV void finish() { vstep = pstep = 0; }
};

// =============================================================================
// filter12 (src/units/filter12.c)
// =============================================================================
template <int CH, bool ADD, bool WIREOUT>
struct Filter12 {
    static constexpr int kWords = 12 + 2 * CH;
    static constexpr bool kUsesFm = false;
    Ramp cutoff, q;
    int lp, bp, hp, f1;
    int d1[CH], d2[CH];
    int f0, df, qstep;

    A2CU_DEV void load(const StatePtr &s, int w) {
        s.ld_ramp(w, cutoff); s.ld_ramp(w + 4, q);
        lp = s.ld(w + 8); bp = s.ld(w + 9); hp = s.ld(w + 10); f1 = s.ld(w + 11);
#pragma unroll
        for (int c = 0; c < CH; ++c) { d1[c] = s.ld(w + 12 + c); d2[c] = s.ld(w + 12 + CH + c); }
        f0 = f1; df = 0; qstep = 0;
    }
    A2CU_DEV void store(const StatePtr &s, int w) const {
        s.st_ramp(w, cutoff); s.st_ramp(w + 4, q);
        s.st(w + 8, lp); s.st(w + 9, bp); s.st(w + 10, hp); s.st(w + 11, f1);
#pragma unroll
        for (int c = 0; c < CH; ++c) { s.st(w + 12 + c, d1[c]); s.st(w + 12 + CH + c, d2[c]); }
    }
    // filter12.c:180-221; arg = transpose
    A2CU_DEV void init(const Ctx &c, int arg, unsigned) {
        ramp_init(cutoff, 0); ramp_init(q, 0);
        write(c, 0, arg, 0, 0);
        write(c, 1, 32768, 0, 0);
        lp = 65536 >> 8; bp = hp = 0;
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) d1[ch] = d2[ch] = 0;
    }
    // filter12.c:141-177. Cooked by the host: cutoff includes transpose, q is
    // the ramper target (32768 or (65536 << 8) / v), lp/bp/hp are >> 8.
    // reg 5 = exact coefficient computed by the host's libm (see engine).
    A2CU_DEV void write(const Ctx &c, int reg, int v, int start, int dur) {
        switch (reg) {
        case 0:
            ramp_set(cutoff, v, start, dur);
            if ((unsigned)dur < 256u) f1 = f12_coeff(c, cutoff.value);
            break;
        case 1: ramp_set(q, v, start, dur); break;
        case 2: lp = v; break;
        case 3: bp = v; break;
        case 4: hp = v; break;
        case 5: f1 = v; break;
        }
    }
    // filter12.c:74-96
    A2CU_DEV void prepare(const Ctx &c, int frames) {
        f0 = f1;
        ramp_prepare(q, frames);
        ramp_prepare(cutoff, frames);
        if (cutoff.delta) {
            ramp_run(cutoff, frames);
            f1 = f12_coeff(c, cutoff.value);
            df = (wsub(f1, f0) + (frames >> 1)) / frames;
        } else
            df = 0;
        qstep = q.delta;
    }
    // filter12.c:97-118
    A2CU_DEV void sample(const Ctx &, int &s0, int &s1, int &o0, int &o1) {
        int f = f0 >> 12;
        int qq = q.value >> 12;
        int in[2] = { s0, s1 };
        int out[2] = { 0, 0 };
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            int d1s = d1[c] >> 4;
            int l = wadd(d2[c], wmul(f, d1s) >> 8);
            int h = wsub(wsub(in[c] >> 5, l), wmul(qq, d1s) >> 8);
            int b = wadd(wmul(f, h >> 4) >> 8, d1[c]);
            out[c] = wadd(wadd(wmul(l, lp), wmul(b, bp)), wmul(h, hp)) >> 3;
            d1[c] = b; d2[c] = l;
        }
        f0 = wadd(f0, df);
        q.value = wadd(q.value, qstep);
        if (WIREOUT) { o0 = wadd(o0, out[0]); if (CH == 2) o1 = wadd(o1, out[1]); }
        else if (ADD) { s0 = wadd(s0, out[0]); if (CH == 2) s1 = wadd(s1, out[1]); }
        else { s0 = out[0]; if (CH == 2) s1 = out[1]; }
    }
    A2CU_DEV bool plain() const { return true; }
    A2CU_DEV void sample_fast(const Ctx &c, int &s0, int &s1, int &o0, int &o1) { sample(c, s0, s1, o0, o1); }
    A2CU_DEV void seed(unsigned) {}
    A2CU_DEV void finish() { qstep = 0; df = 0; }
};

// =============================================================================
// waveshaper (src/units/waveshaper.c:55-108)
// =============================================================================
template <int CH, bool ADD, bool WIREOUT>
struct WaveShaper {
    static constexpr int kWords = 4;
    static constexpr bool kUsesFm = false;
    Ramp amount;
    int step;
    A2CU_DEV void load(const StatePtr &s, int w) { s.ld_ramp(w, amount); step = 0; }
    A2CU_DEV void store(const StatePtr &s, int w) const { s.st_ramp(w, amount); }
    A2CU_DEV void init(const Ctx &, int, unsigned) { ramp_init(amount, 0); }
    A2CU_DEV void write(const Ctx &, int, int v, int start, int dur) { ramp_set(amount, v, start, dur); }
    A2CU_DEV void prepare(const Ctx &, int frames) { ramp_prepare(amount, frames); step = amount.delta; }
    A2CU_DEV int shape(int v, int a, int a3p1, int asqr) {
        int vsqr = (int)(((long long)v * v) >> 22);
        long long vout = (long long)v * a3p1;
        long long sqrsub = (long long)a * vsqr;
        if (v >= 0) vout -= sqrsub; else vout += sqrsub;
        vout /= (((long long)asqr * vsqr) >> 16) + (1 << 24);
        return (int)vout;
    }
    A2CU_DEV void sample(const Ctx &, int &s0, int &s1, int &o0, int &o1) {
        int a = amount.value;
        int a3p1 = wadd(wadd((int)((unsigned)a << 1), a), 1 << 24);
        int asqr = (int)(((long long)(a >> 4) * (a >> 4)) >> 24);
        int r0 = shape(s0, a, a3p1, asqr);
        int r1 = CH == 2 ? shape(s1, a, a3p1, asqr) : 0;
        amount.value = wadd(amount.value, step);
        if (WIREOUT) { o0 = wadd(o0, r0); if (CH == 2) o1 = wadd(o1, r1); }
        else if (ADD) { s0 = wadd(s0, r0); if (CH == 2) s1 = wadd(s1, r1); }
        else { s0 = r0; if (CH == 2) s1 = r1; }
    }
    A2CU_DEV bool plain() const { return true; }
    A2CU_DEV void sample_fast(const Ctx &c, int &s0, int &s1, int &o0, int &o1) { sample(c, s0, s1, o0, o1); }
    A2CU_DEV void seed(unsigned) {}
    A2CU_DEV void finish() { step = 0; }
};

// =============================================================================
// dcblock (src/units/dcblock.c:65-94): the filter12 topology with a fixed high-pass output.
// The coefficient (dcb_pitch2coeff, :57-63, host libm) arrives cooked from the host.
// =============================================================================
template <int CH, bool ADD, bool WIREOUT>
struct DcBlock {
    static constexpr int kWords = 1 + 2 * CH;
    static constexpr bool kUsesFm = false;
    int f1;
    int d1[CH], d2[CH];
    A2CU_DEV void load(const StatePtr &s, int w) {
        f1 = s.ld(w);
#pragma unroll
        for (int c = 0; c < CH; ++c) { d1[c] = s.ld(w + 1 + c); d2[c] = s.ld(w + 1 + CH + c); }
    }
    A2CU_DEV void store(const StatePtr &s, int w) const {
        s.st(w, f1);
#pragma unroll
        for (int c = 0; c < CH; ++c) { s.st(w + 1 + c, d1[c]); s.st(w + 1 + CH + c, d2[c]); }
    }
    // dcblock.c:117-146; the default cutoff's coefficient follows as a cooked write
    A2CU_DEV void init(const Ctx &, int, unsigned) {
        f1 = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c) d1[c] = d2[c] = 0;
    }
    A2CU_DEV void write(const Ctx &, int reg, int v, int, int) { if (reg == 0) f1 = v; }
    A2CU_DEV void prepare(const Ctx &, int) {}
    A2CU_DEV void sample(const Ctx &, int &s0, int &s1, int &o0, int &o1) {
        const int f = f1 >> 12;
        int in[2] = { s0, s1 };
        int out[2] = { 0, 0 };
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int d1s = d1[c] >> 4;
            const int l = wadd(d2[c], wmul(f, d1s) >> 8);
            const int h = wsub(wsub(in[c] >> 5, l), (int)((unsigned)d1s << 4));
            const int b = wadd(wmul(f, h >> 4) >> 8, d1[c]);
            out[c] = (int)((unsigned)h << 5);
            d1[c] = b; d2[c] = l;
        }
        if (WIREOUT) { o0 = wadd(o0, out[0]); if (CH == 2) o1 = wadd(o1, out[1]); }
        else if (ADD) { s0 = wadd(s0, out[0]); if (CH == 2) s1 = wadd(s1, out[1]); }
        else { s0 = out[0]; if (CH == 2) s1 = out[1]; }
    }
    A2CU_DEV bool plain() const { return true; }
    A2CU_DEV void sample_fast(const Ctx &c, int &s0, int &s1, int &o0, int &o1) { sample(c, s0, s1, o0, o1); }
    A2CU_DEV void seed(unsigned) {}
    A2CU_DEV void finish() {}
};

// =============================================================================
// limiter (src/units/limiter.c:51-153). release and threshold arrive cooked
// ((v << 8) / samplerate and max(v << 8, 256), :187-199); the peak follower's unsigned
// arithmetic (peak -= release may wrap) is kept as written.
// =============================================================================
template <int CH, bool ADD, bool WIREOUT>
struct Limiter {
    static constexpr int kWords = 3;
    static constexpr bool kUsesFm = false;
    unsigned threshold, peak;
    int release;
    A2CU_DEV void load(const StatePtr &s, int w) {
        threshold = (unsigned)s.ld(w); release = s.ld(w + 1); peak = (unsigned)s.ld(w + 2);
    }
    A2CU_DEV void store(const StatePtr &s, int w) const {
        s.st(w, (int)threshold); s.st(w + 1, release); s.st(w + 2, (int)peak);
    }
    // limiter.c:156-184: arg = samplerate-cooked default release
    A2CU_DEV void init(const Ctx &, int arg, unsigned) {
        release = arg;
        threshold = (unsigned)((1 << 16) << 8);
        peak = 32768u << 8;
    }
    A2CU_DEV void write(const Ctx &, int reg, int v, int, int) {
        if (reg == 0) release = v; else if (reg == 1) threshold = (unsigned)v;
    }
    A2CU_DEV void prepare(const Ctx &, int) {}
    A2CU_DEV static unsigned uabs(int x) { return x < 0 ? 0u - (unsigned)x : (unsigned)x; }
    A2CU_DEV void sample(const Ctx &, int &s0, int &s1, int &o0, int &o1) {
        unsigned p;
        if (CH == 1) p = uabs(s0);
        else {
            // limiter.c:109-113: int abs values, unsigned combination
            const int lp = (int)uabs(s0), rp = (int)uabs(s1);
            p = (unsigned)(lp > rp ? lp : rp);
            p = p + ((p - uabs(wsub(lp, rp))) >> 1);
        }
        if (p > peak) peak = p;
        else {
            peak -= (unsigned)release;
            if (peak < threshold) peak = threshold;
            p = peak;
        }
        const int gain = (int)((32767LL << 16) / (long long)((p + 511u) >> 9));
        const int r0 = mulshr(s0, gain, 16);
        const int r1 = CH == 2 ? mulshr(s1, gain, 16) : 0;
        if (WIREOUT) { o0 = wadd(o0, r0); if (CH == 2) o1 = wadd(o1, r1); }
        else if (ADD) { s0 = wadd(s0, r0); if (CH == 2) s1 = wadd(s1, r1); }
        else { s0 = r0; if (CH == 2) s1 = r1; }
    }
    A2CU_DEV bool plain() const { return true; }
    A2CU_DEV void sample_fast(const Ctx &c, int &s0, int &s1, int &o0, int &o1) { sample(c, s0, s1, o0, o1); }
    A2CU_DEV void seed(unsigned) {}
    A2CU_DEV void finish() {}
};

// =============================================================================
// dc (src/units/dc.c:57-140): a ramped (LINEAR) or stepped (STEP, one interpolated
// "transient" sample at the switch point) constant on 1 or 2 outputs.
// =============================================================================
template <int NOUT, bool ADD, bool WIREOUT>
struct Dc {
    static constexpr int kWords = 5;
    static constexpr bool kUsesFm = false;
    Ramp value;
    int mode;               // 0 STEP, 1 LINEAR (dc.c:34-43)
    // segment-local (STEP): frames [0, fill_end) hold the old value, frame trans_at the transient
    int k, fill_end, trans_at, tv, old_value;
    A2CU_DEV void load(const StatePtr &s, int w) { s.ld_ramp(w, value); mode = s.ld(w + 4); k = 0; fill_end = 0; trans_at = -1; tv = 0; old_value = 0; }
    A2CU_DEV void store(const StatePtr &s, int w) const { s.st_ramp(w, value); s.st(w + 4, mode); }
    A2CU_DEV void init(const Ctx &, int, unsigned) { ramp_init(value, 0); mode = 1; }   // dc.c:153-180
    // dc.c:183-247
    A2CU_DEV void write(const Ctx &, int reg, int v, int start, int dur) {
        if (reg == 1) {
            mode = v >> 16;
            if (mode != 0 && mode != 1) mode = 0;
            return;
        }
        if (mode == 0) {
            value.target = (int)((unsigned)v << 8);
            value.timer = (int)(((unsigned)dur >> 1) - (unsigned)start);
            if (value.timer <= 0) { value.value = value.target; value.timer = 0; }
        } else
            ramp_set(value, v, start, dur);
    }
    A2CU_DEV void prepare(const Ctx &, int frames) {
        k = 0; fill_end = 0; trans_at = -1;
        if (mode == 1) { ramp_prepare(value, frames); return; }
        old_value = value.value;
        int s = 0;
        if (value.timer >= 256) {                           // dc.c:72-92
            if ((value.timer >> 8) >= frames) { s = frames; value.timer -= frames << 8; }
            else { s = value.timer >> 8; value.timer &= 0xff; }
            fill_end = s;
        }
        if (value.timer < 256 && s < frames) {              // dc.c:94-108
            tv = wadd(wmul(value.value >> 4, value.timer), wmul(value.target >> 4, 256 - value.timer)) >> 4;
            trans_at = s;
            value.timer = 0;
            value.value = value.target;
        }
    }
    A2CU_DEV void sample(const Ctx &, int &s0, int &s1, int &o0, int &o1) {
        int v;
        if (mode == 1) { v = value.value; value.value = wadd(value.value, value.delta); }
        else v = k < fill_end ? old_value : (k == trans_at ? tv : value.target);
        ++k;
        if (WIREOUT) { o0 = wadd(o0, v); if (NOUT == 2) o1 = wadd(o1, v); }
        else if (ADD) { s0 = wadd(s0, v); if (NOUT == 2) s1 = wadd(s1, v); }
        else { s0 = v; if (NOUT == 2) s1 = v; }
    }
    A2CU_DEV bool plain() const { return true; }
    A2CU_DEV void sample_fast(const Ctx &c, int &s0, int &s1, int &o0, int &o1) { sample(c, s0, s1, o0, o1); }
    A2CU_DEV void seed(unsigned) {}
    A2CU_DEV void finish() {}
};

// =============================================================================
// fm1..fm4r (src/units/fm.c). NOPS operators, OSBITS oversampling bits,
// PAR 0 = chain, 1 = parallel modulators, 2 = ring modulator pair.
// As built, fm.c does not see A2_HIFI: fm1 1x, fm2/fm2r 2x, the rest 4x.
// =============================================================================
struct FmOp {
    Ramp a, fb, p;
    int last_pitch;
    unsigned phase, dphase;
    int last;
};

template <int NOPS, int OSBITS, int PAR, bool ADD, bool WIREOUT>
struct Fm {
    static constexpr int kWords = 16 * NOPS;
    static constexpr bool kUsesFm = true;
    FmOp op[NOPS];
    int astep[NOPS], fbstep[NOPS];

    A2CU_DEV void load(const StatePtr &s, int w) {
#pragma unroll
        for (int i = 0; i < NOPS; ++i) {
            int b = w + 16 * i;
            s.ld_ramp(b, op[i].a); s.ld_ramp(b + 4, op[i].fb); s.ld_ramp(b + 8, op[i].p);
            op[i].last_pitch = s.ld(b + 12);
            op[i].phase = (unsigned)s.ld(b + 13);
            op[i].dphase = (unsigned)s.ld(b + 14);
            op[i].last = s.ld(b + 15);
            astep[i] = fbstep[i] = 0;
        }
    }
    A2CU_DEV void store(const StatePtr &s, int w) const {
#pragma unroll
        for (int i = 0; i < NOPS; ++i) {
            int b = w + 16 * i;
            s.st_ramp(b, op[i].a); s.st_ramp(b + 4, op[i].fb); s.st_ramp(b + 8, op[i].p);
            s.st(b + 12, op[i].last_pitch);
            s.st(b + 13, (int)op[i].phase);
            s.st(b + 14, (int)op[i].dphase);
            s.st(b + 15, op[i].last);
        }
    }
    // fm.c:328-336
    A2CU_DEV void set_phase(int ph, unsigned sst) {
#pragma unroll
        for (int i = 0; i < NOPS; ++i) {
            int ssph = wadd(ph, (int)((sst * (op[i].dphase >> 8)) >> 8));
            op[i].phase = (unsigned)(wmul(ssph, 2048) >> 8);
        }
    }
    // fm.c:339-408; arg = transpose + basepitch
    A2CU_DEV void init(const Ctx &c, int arg, unsigned sst) {
#pragma unroll
        for (int i = 0; i < NOPS; ++i) {
            ramp_init(op[i].a, 0); ramp_init(op[i].fb, 0); ramp_init(op[i].p, arg);
            op[i].last_pitch = 0; op[i].last = 0;
        }
        op[0].dphase = p2i(c.ptab, op[0].p.value >> 8);
#pragma unroll
        for (int i = 1; i < NOPS; ++i) op[i].dphase = op[0].dphase;
        set_phase(0, sst);
    }
    // fm.c:411-483; op0 pitch is cooked (incl. transpose + basepitch)
    A2CU_DEV void write(const Ctx &, int reg, int v, int start, int dur) {
        if (reg == 0) { set_phase(v, (unsigned)start); return; }
        int o = (reg - 1) / 3, which = (reg - 1) % 3;
#pragma unroll
        for (int i = 0; i < NOPS; ++i)
            if (i == o) {
                if (which == 0) ramp_set(op[i].p, v, start, dur);
                else if (which == 1) ramp_set(op[i].a, v, start, dur);
                else ramp_set(op[i].fb, v, start, dur);
            }
    }
    // fm.c:194-210 with fm_run_pitch :125-140
    A2CU_DEV void prepare(const Ctx &c, int frames) {
        int detune = 0;
#pragma unroll
        for (int i = 0; i < NOPS; ++i) {
            ramp_prepare(op[i].a, frames);
            ramp_prepare(op[i].fb, frames);
            ramp_prepare(op[i].p, frames);
            ramp_run(op[i].p, frames >> 1);
            int np = wadd(op[i].p.value, detune) >> 8;
            if (np != op[i].last_pitch) {
                op[i].dphase = p2i(c.ptab, np);
                op[i].last_pitch = np;
            }
            detune = op[0].p.value;
            astep[i] = op[i].a.delta; fbstep[i] = op[i].fb.delta;
        }
    }
    // fm.c:111-122
    A2CU_DEV int osc(const Ctx &c, FmOp &o, int mod) {
        int fb = mulshr(o.last, o.fb.value, 17);
        unsigned ph = (o.phase + (unsigned)mod + (unsigned)fb) >> 5;
        o.last = lerp16_packed(c.fmsine, ph & ((2048u << 8) - 1));
        return mulshr(o.last, o.a.value, 16);
    }
    // fm.c:150-163 / :170-192
    A2CU_DEV int subsample(const Ctx &c) {
        if (PAR == 2) {
            int v[2];
            if (NOPS == 2) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    v[i] = osc(c, op[i], 0);
                    op[i].phase += op[i].dphase >> OSBITS;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    constexpr int dummy = 0; (void)dummy;
                    int m = osc(c, op[(i + 2) % NOPS], 0);
                    v[i] = osc(c, op[i], m);
                    op[i].phase += op[i].dphase >> OSBITS;
                    op[(i + 2) % NOPS].phase += op[(i + 2) % NOPS].dphase >> OSBITS;
                }
            }
            return (int)(((long long)v[0] * v[1]) >> 23);
        }
        int v = 0;
#pragma unroll
        for (int i = NOPS - 1; i >= 0; --i) {
            if (i && PAR == 1) v = wadd(v, osc(c, op[i], 0));
            else v = osc(c, op[i], v);
            op[i].phase += op[i].dphase >> OSBITS;
        }
        return v;
    }
    // fm.c:211-232
    A2CU_DEV void sample(const Ctx &c, int &s0, int &s1, int &o0, int &o1) {
        int vsum = 0;
#pragma unroll
        for (int os = 0; os < (1 << OSBITS); ++os) vsum = wadd(vsum, subsample(c));
#pragma unroll
        for (int i = 0; i < NOPS; ++i) {
            op[i].a.value = wadd(op[i].a.value, astep[i]);
            op[i].fb.value = wadd(op[i].fb.value, fbstep[i]);
            op[i].phase += op[i].dphase & ((1u << OSBITS) - 1);
        }
        int v = vsum >> OSBITS;
        if (WIREOUT) o0 = wadd(o0, v);
        else if (ADD) s0 = wadd(s0, v);
        else s0 = v;
    }
    A2CU_DEV void seed(unsigned) {}
    A2CU_DEV bool plain() const { return true; }
    A2CU_DEV void sample_fast(const Ctx &c, int &s0, int &s1, int &o0, int &o1) { sample(c, s0, s1, o0, o1); }
    A2CU_DEV void finish() {
#pragma unroll
        for (int i = 0; i < NOPS; ++i) astep[i] = fbstep[i] = 0;
    }
};

// =============================================================================
// Chain: compile-time list of units
// =============================================================================
template <class... Us> struct Chain;

template <> struct Chain<> {
    static constexpr int kWords = 0;
    static constexpr bool kUsesFm = false;
    A2CU_DEV void load(const StatePtr &, int) {}
    A2CU_DEV void store(const StatePtr &, int) const {}
    A2CU_DEV void init_unit(const Ctx &, int, int, unsigned) {}
    A2CU_DEV void write(const Ctx &, int, int, int, int, int) {}
    A2CU_DEV void seed_unit(int, unsigned) {}
    A2CU_DEV void prepare(const Ctx &, int) {}
    A2CU_DEV void sample(const Ctx &, int &, int &, int &, int &) {}
    A2CU_DEV void sample_fast(const Ctx &, int &, int &, int &, int &) {}
    A2CU_DEV bool plain() const { return true; }
    A2CU_DEV void finish() {}
};

template <class U, class... Rest> struct Chain<U, Rest...> {
    static constexpr int kWords = U::kWords + Chain<Rest...>::kWords;
    static constexpr bool kUsesFm = U::kUsesFm || Chain<Rest...>::kUsesFm;
    U u;
    Chain<Rest...> rest;
    A2CU_DEV void load(const StatePtr &s, int w) { u.load(s, w); rest.load(s, w + U::kWords); }
    A2CU_DEV void store(const StatePtr &s, int w) const { u.store(s, w); rest.store(s, w + U::kWords); }
    A2CU_DEV void init_unit(const Ctx &c, int unit, int arg, unsigned sst) {
        if (unit == 0) u.init(c, arg, sst); else rest.init_unit(c, unit - 1, arg, sst);
    }
    A2CU_DEV void write(const Ctx &c, int unit, int reg, int v, int start, int dur) {
        if (unit == 0) u.write(c, reg, v, start, dur); else rest.write(c, unit - 1, reg, v, start, dur);
    }
    A2CU_DEV void seed_unit(int unit, unsigned sd) {
        if (unit == 0) u.seed(sd); else rest.seed_unit(unit - 1, sd);
    }
    A2CU_DEV void prepare(const Ctx &c, int frames) { u.prepare(c, frames); rest.prepare(c, frames); }
    A2CU_DEV void sample(const Ctx &c, int &s0, int &s1, int &o0, int &o1) {
        u.sample(c, s0, s1, o0, o1); rest.sample(c, s0, s1, o0, o1);
    }
    A2CU_DEV void sample_fast(const Ctx &c, int &s0, int &s1, int &o0, int &o1) {
        u.sample_fast(c, s0, s1, o0, o1); rest.sample_fast(c, s0, s1, o0, o1);
    }
    A2CU_DEV bool plain() const { return u.plain() && rest.plain(); }
    A2CU_DEV void finish() { u.finish(); rest.finish(); }
};

}  // namespace a2cu
