// a2cu_engine.cu - host side of the voice engine and the C ABI of include/a2cu.h.
//
// The host owns: wave preparation (waves.c:59-151 semantics), the pitch and FM
// sine tables (built with the host libm exactly like pitch.c:70-96 and
// fm.c:486-501, then uploaded - never recomputed with device libm), event
// queues, and the launch sequence per window:
//
//   H2D event CSR -> memset bus accumulators -> render_bank<Chain> per bank
//   -> mix_groups -> mix_root -> (D2H master)
//
// No CPU fallback exists: every sample is produced by the kernels.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/a2cu.h"
#include "a2cu_kernels.cuh"
#include "a2cu_bus.cuh"
#include "a2cu_waves.cuh"
#include "a2cu_registry.h"

using namespace a2cu;

static thread_local char g_err[512] = "";

// Optional host-side statistics of block mode (env A2CU_STATS=1, printed at close)
#include <time.h>
static inline double now_us() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
struct BlockStats { double flush_us = 0, download_us = 0, sync_us = 0, begin_us = 0; long flushes = 0, downloads = 0, begins = 0, procs = 0; };
struct WindowStats { double prep_us = 0, stage_us = 0, api_us = 0, mix_us = 0; long windows = 0; };
static int fail(int code, const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
#define CK(call)                                                              \
    do {                                                                      \
        cudaError_t err__ = (call);                                           \
        if (err__ != cudaSuccess)                                             \
            return fail(A2CU_ECUDA, #call ": %s", cudaGetErrorString(err__)); \
    } while (0)

// cudaMemset runs on the legacy default stream and may return before the device is done. The engine's
// kernels run on e->stream, which a caller may have set to a NON-BLOCKING stream (torch.cuda.Stream is
// one): nothing then orders a creation-time clear before the first render kernel. Every such clear is
// therefore completed before the call returns (creation paths only, never per window).
static cudaError_t memcpy_done(void *dst, const void *src, size_t n, cudaMemcpyKind kind) {
    cudaError_t r = cudaMemcpy(dst, src, n, kind);      // (H2D from pageable memory returns once staged, D2D at once)
    if (r == cudaSuccess) r = cudaStreamSynchronize(0);
    return r;
}
static cudaError_t memset_done(void *p, int v, size_t n) {
    cudaError_t r = cudaMemset(p, v, n);
    if (r == cudaSuccess) r = cudaStreamSynchronize(0);
    return r;
}

// ---------------------------------------------------------------------------
// Chain registry: signature string -> kernel. The kernels are instantiated in
// their own translation units (a2cu_reg_*.cu, compiled in parallel by build.py).
// ---------------------------------------------------------------------------
std::map<std::string, KernelEntry> &a2cu_registry() {
    static std::map<std::string, KernelEntry> r;
    return r;
}
static std::map<std::string, KernelEntry> &registry() { return a2cu_registry(); }
static void register_all() {
    static bool done = false;
    if (done) return;
    done = true;
    a2cu_register_bank_wt();
    a2cu_register_bank_fm();
}
static void register_split() { a2cu_register_split(); }

// ---------------------------------------------------------------------------
// Host-side tables (same libm calls as the reference, uploaded as integers)
// ---------------------------------------------------------------------------
struct HostTables {
    unsigned ptab[128];     // {base, coeff} x 64, pitch.c:70-96
    int16_t fmsine[2049];   // fm.c:486-501
    int fmsine_packed[2048];    // {sine[i] | (sine[i + 1] - sine[i]) << 16}: what the kernels read (lerp16_packed)
    HostTables() {
        unsigned b = 0x80000000u;
        for (unsigned i = 0; i < 64; ++i) {
            unsigned b2 = (unsigned)(long long)((double)0x80000000u * powf(2.0f, (i + 1) * (1.0f / 64)) + 0.5f);
            ptab[2 * i] = b;
            ptab[2 * i + 1] = (b2 - b + 128) >> 8;
            b = b2;
        }
        for (int s = 0; s < 2049; ++s) fmsine[s] = (int16_t)(sin(s * 2.0f * M_PI / 2048) * 32767.0f);
        for (int s = 0; s < 2048; ++s)
            fmsine_packed[s] = (int)((unsigned)(uint16_t)fmsine[s] | ((unsigned)(fmsine[s + 1] - fmsine[s]) << 16));
    }
    unsigned p2i(int pitch) const {     // pitch.c:57-67 (shift count & 31 as on x86-64)
        int n = pitch & 0xffff;
        int oct = pitch >> 16;
        unsigned dph = ptab[2 * (n >> 10) + 1] * (unsigned)(n & 0x3ff);
        dph >>= 2;
        dph += ptab[2 * (n >> 10)];
        return dph >> ((7 - oct) & 31);
    }
    int f12_coeff(int cutoff_value, int samplerate) const {    // filter12.c:65-72
        float f = p2i(cutoff_value >> 8) * (261.626f / 16777216.0f);
        if (f > (samplerate >> 2)) return 362 << 16;
        return (int)(512.0f * 65536.0f * sin(M_PI * f / samplerate));
    }
};
static const HostTables &tables() {
    static HostTables t;
    return t;
}

// f12_pitch2coeff (filter12.c:65-72) for every argument it can see at one sample rate: the 16
// fraction bits of the pitch x the 32 values of (7 - octave) & 31 (a2_P2I, pitch.c:57-67).
// Evaluated with the HOST libm, the same expression as the reference, once per sample rate and
// process (~2 M sin calls); the device only looks results up (a2cu_device.cuh f12_coeff), so
// a ramping cutoff is bit-exact by construction.
#include <mutex>
static const std::vector<int> &f12_table(int samplerate) {
    static std::mutex mu;
    static std::map<int, std::vector<int>> memo;
    std::lock_guard<std::mutex> lock(mu);
    std::vector<int> &t = memo[samplerate];
    if (t.empty()) {
        t.resize((size_t)32 << 16);
        const HostTables &h = tables();
        for (int shift = 0; shift < 32; ++shift)
            for (int n = 0; n < 65536; ++n) {
                // a pitch whose a2_P2I shift count is `shift`: octave 7 - shift
                const int pitch = ((7 - shift) << 16) | n;
                t[((size_t)shift << 16) | n] = h.f12_coeff((int)((unsigned)pitch << 8), samplerate);
            }
    }
    return t;
}


// ---------------------------------------------------------------------------
// Noise planner: host-side control-rate twin of the wtosc unit.
//
// The noise wave draws from ONE LCG shared by every voice (and the VM's RAND)
// in tree-walk order (wtosc.c:135-144, core.c:1401-1409), so the start state of
// each noise segment depends on how many draws every earlier voice made. Voices
// that select the noise wave get a host "mirror" of their oscillators' control
// state (pitch ramper, dphase, phase, mode) - created by reading the device
// state back once - which replays exactly the per-Process()-call logic of
// a2cu_device.cuh WtOsc::prepare/finish with the sample loop in closed form.
// It yields the draw count per segment; the LCG start state is delivered to the
// device as an EV_SEED record. No audio is computed on the host.
// ---------------------------------------------------------------------------
struct HRamp { int value, target, delta, timer; };
static inline int h_add(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
static inline int h_sub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
static inline int h_mul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
static void hramp_prepare(HRamp &r, int frames) {      // a2_dsp.h:128-149
    if (!r.timer) { r.value = r.target; r.delta = 0; }
    else if (frames <= (r.timer >> 8)) {
        r.delta = (int)(((long long)h_sub(r.target, r.value) << 8) / r.timer);
        r.timer -= frames << 8;
    } else { r.delta = h_sub(r.target, r.value) / frames; r.timer = 0; }
}
static void hramp_run(HRamp &r, int frames) { r.value = h_add(r.value, h_mul(r.delta, frames)); }
static void hramp_set(HRamp &r, int target, int start, int dur) {   // a2_dsp.h:161-170
    r.target = (int)((unsigned)target << 8);
    r.timer = dur + start;
    if (r.timer < 256) r.value = r.target;
    else r.value = h_add(r.value, h_mul(r.delta, start) >> 8);
}

struct OscMirror {
    HRamp p;
    int p_ramping;
    unsigned dphase;
    unsigned long long phase;
    int wave, mode;     // OscMode

    void load(const int *w) {       // word layout of WtOsc::store
        p.value = w[0]; p.target = w[1]; p.delta = w[2]; p.timer = w[3];
        dphase = (unsigned)w[8];
        phase = (unsigned)w[9] | ((unsigned long long)(unsigned)w[10] << 32);
        wave = w[12] >> 8; mode = w[12] & 0xff;
        p_ramping = w[13];
    }
    void set_phase(const a2cu_engine *e, int phv, unsigned sst);
    void init(const a2cu_engine *e, int arg, unsigned sst) {
        wave = -1;
        p.value = p.target = (int)((unsigned)arg << 8); p.delta = p.timer = 0;
        dphase = tables().p2i(p.value >> 8);
        p_ramping = 0;
        set_phase(e, 0, sst);
        mode = OSC_OFF;
    }
    void write(const a2cu_engine *e, int reg, int v, int start, int dur);
    void run_pitch(int frames) {        // wtosc.c:89-105
        hramp_prepare(p, frames);
        if (dphase && (!p.timer && !p_ramping)) return;
        unsigned lastv = (unsigned)p.value;
        hramp_run(p, frames);
        p_ramping = p.delta;
        dphase = tables().p2i((int)((lastv + (unsigned)p.value) >> 9));
    }
    // One Process() call: returns the number of LCG draws (noise mode) and
    // whether the segment is a noise segment at all.
    int segment(const a2cu_engine *e, int frames, bool *is_noise);
};

struct VoiceMirror {
    int alive = 0;
    std::vector<std::pair<int, OscMirror>> osc;     // (unit index, twin)
    size_t cursor = 0;                              // planner scratch
};

// ---------------------------------------------------------------------------
// Engine objects
// ---------------------------------------------------------------------------
// A wave as the host knows it: geometry only. The samples live in the device pool; pads, mip
// levels and Hermite coefficients are produced there (a2cu_waves.cuh).
struct HostWave {
    int type;
    unsigned flags, period;
    unsigned size[kMipLevels];      // excluding pads
    size_t first[kMipLevels];       // pool index of the level's FIRST PAD sample
    size_t span[kMipLevels];        // samples incl. pads (0: level absent)
    int coff[kMipLevels];           // coefficient-pool index of sample 0, -1: none
    std::string name;
    int cbegin = -1, ccount = 0;    // range in the Hermite coefficient pool
    size_t total() const { size_t t = 0; for (int l = 0; l < kMipLevels; ++l) t += span[l]; return t; }
};

struct HostEvent {
    uint64_t time;
    uint32_t seq;
    int voice;
    uint32_t y;     // kind | unit << 8 | reg << 16
    int value;
    uint32_t dur;
};

struct Bank {
    std::vector<a2cu_unitspec> chain;
    KernelEntry k;
    int nvoices = 0;
    size_t stride = 0;
    int *d_state = nullptr;
    int *d_bus = nullptr;
    unsigned *d_noise = nullptr;
    std::vector<int> transpose, group;
    std::vector<HostEvent> events;
    // a2cu_bank_write_all of a register whose cooked value does not depend on the voice's transpose:
    // kept as ONE record (+ the value column) and expanded straight into the staging buffer
    struct BulkEvent {
        uint64_t time; uint32_t seq, y, dur; int32_t scalar; bool has_values; std::vector<int32_t> values;
    };
    std::vector<BulkEvent> bulk;
    uint32_t seq = 0;
    // per-window device buffers (grown on demand)
    unsigned *d_evoff = nullptr;    // views into d_ev (bank mode): CSR offsets, then records
    uint4 *d_evrecs = nullptr;
    uint4 *d_ev = nullptr;
    size_t ev_cap = 0;
    bool has_noise = false;
    bool exotic = false;    // selected a noise / one-shot sampled wave: render_bank only
    bool raw_taps = false;  // selected a wave without a coefficient table: render_split's raw-tap variant
    // a structure without a fused kernel: render_generic (a2cu_bus.cuh) with a scratch row per voice
    bool generic = false;
    GenericChain gchain;
    int *d_scratch = nullptr;
    cudaEvent_t ev_consumed = nullptr;  // recorded after the last kernel that reads d_ev (bank mode)
    bool enabled = true;    // a2cu_bank_enable: a disabled bank is paused (not rendered, events kept)
    int stage_wave = -1;    // wave whose coefficient table render_split stages in shared memory
    uint32_t stamp = 0;
    // drop-in ("block") mode: dynamic slots + per-flush recording
    bool dynamic = false;
    std::vector<int> free_slots, deferred_free;   // freed slots are reusable from the next block
    int used = 0;
    std::vector<uint4> bev;
    std::vector<VoiceRun> bruns;
    int cur_slot = -1;
    std::vector<uint32_t> slot_mark;    // flush id in which a slot got its first run
    uint32_t flush_id = 1;
    bool dup_runs = false;              // some slot has more than one run in this flush
    VoiceRun *d_runs = nullptr;
    size_t runs_cap = 0;
    // slot has an entry in a2cu_engine::mirrors (spares the map lookup on the
    // per-call recording path)
    std::vector<uint8_t> has_mirror;
    bool mirrored(int slot) const { return (size_t)slot < has_mirror.size() && has_mirror[slot]; }
    void set_mirrored(int slot, bool on) {
        if ((size_t)slot >= has_mirror.size()) { if (!on) return; has_mirror.resize((size_t)slot + 256, 0); }
        has_mirror[slot] = on ? 1 : 0;
    }
};

struct MixHostEvent {
    uint64_t time;
    uint32_t seq;
    int target, reg, value;
    uint32_t dur;
};

struct a2cu_engine {
    // host-side statistics (printed at close with A2CU_STATS=1); per engine: different states may
    // run on different threads (audiality2.h.cmake:163-166)
    BlockStats bs;
    WindowStats ws;
    // environment toggles, read once in a2cu_open (A/B switches for profiles/)
    bool env_stats = false, env_no_copy_stream = false, env_no_fuse = false, env_no_stage = false, env_one_set = false,
         env_dry = false,     // A2CU_DRY: block mode records but launches nothing (measures the recording cost; silence)
         env_no_side_streams = false, env_split_always = false;
    int sm_count = 148;
    int device = 0, samplerate = 48000, channels = 2;
    int basepitch = 0;
    uint32_t msdur = 0;
    uint64_t now = 0;           // 24:8
    uint32_t root_wake = 1000000;
    uint32_t noiseseed = 324357;
    cudaStream_t stream = 0;
    bool post_root = true;
    int out_fmt = 0;            // a2cu_set_output_format
    std::vector<HostWave> waves;
    bool waves_dirty = true;
    WaveDesc *d_waves = nullptr;
    int16_t *d_pool = nullptr;
    int4 *d_cpool = nullptr;
    size_t pool_cap = 0, waves_cap = 0, cpool_cap = 0;
    size_t pool_used = 0, cpool_used = 0;   // device arenas: waves are appended, never moved
    unsigned *d_ptab = nullptr;
    int *d_fmsine = nullptr;
    int *d_f12tab = nullptr;        // f12_table(samplerate)
    std::vector<Bank *> banks;
    std::vector<RenderParams> params;
    int ngroups = 0;
    int *d_gstate = nullptr;
    int gstate_cap = 0;
    int *d_rstate = nullptr;
    std::vector<MixHostEvent> mixev;
    uint32_t mixseq = 0;
    // root panmix: a write / ramp is (or may still be) in flight until this time; the windows it
    // touches, and one more (the ramper snaps value = target one segment later, a2_dsp.h:130-134),
    // take mix_root's single-CTA general path
    uint64_t root_until = 0;
    bool root_written = false, root_extra = false;
    MixEvent *d_mixev = nullptr;
    size_t mixev_cap = 0;
    int *d_acc = nullptr;
    size_t acc_cap = 0;
    size_t acc_clean_n = 0;     // d_acc is all-zero for this layout (see run_window)
    unsigned *d_fuse_counter = nullptr;     // render_split's fused root stage: CTA tickets
    bool fused_root = false;                // last window ran the root stage inside render_split
    int acc_clean_W = 0;
    int *d_master = nullptr;
    size_t master_cap = 0;
    // pinned staging
    void *h_stage = nullptr;
    size_t stage_cap = 0;
    // staging ring: run_window alternates between kStageRing pinned buffers so a
    // window can be staged while the previous one's H2D copies are still queued
    static const int kStageRing = 3;
    void *stage_buf[kStageRing] = {nullptr, nullptr, nullptr};
    size_t stage_bufcap[kStageRing] = {0, 0, 0};
    cudaEvent_t stage_done[kStageRing] = {nullptr, nullptr, nullptr};
    bool stage_used[kStageRing] = {false, false, false};
    int stage_pos = 0;
    // event uploads go through their own stream so that the H2D copy of window i+1 overlaps the
    // kernels of window i (it only waits for the last kernel that read the same bank's event buffer)
    cudaStream_t copy_stream = nullptr;
    bool copy_used = false;         // this window's uploads went through copy_stream
    // Banks of one window are independent until the bus stage (they only add into the buses with
    // integer atomics): a window with several banks launches them on side streams, forked from and
    // joined back into the engine's stream, so that small banks share the chip instead of queueing.
    static const int kSideStreams = 3;
    cudaStream_t side[kSideStreams] = {nullptr, nullptr, nullptr};
    cudaEvent_t side_fork = nullptr, side_join[kSideStreams] = {nullptr, nullptr, nullptr};
    // pipelined API (a2cu_submit / a2cu_collect): result slots
    static const int kSlots = 4;
    struct Slot {
        cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, done = nullptr;
        int32_t *h_out = nullptr;
        size_t cap = 0, n = 0;
        bool busy = false;
    } slots[kSlots];
    int slot_pos = 0;
    int32_t *h_out = nullptr;
    size_t hout_cap = 0;
    uint64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0;
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    float last_ms = 0.f, last_mix_ms = 0.f;
    // last window (for a2cu_apply_root_stage)
    MixParams last_mix;
    // noise planner
    std::map<uint64_t, VoiceMirror> mirrors;    // key: bank << 32 | slot
    uint32_t *noise_ptr = nullptr;              // shared LCG (host's st->noisestate in drop-in mode)
    uint32_t stamp = 0;                         // creation order (tree-walk order is newest first)
    bool use_split = true;                      // allow render_split where eligible
    bool noise_seen = false;                    // some voice selected a noise wave (planner armed)
    unsigned long long *d_prof = nullptr;       // render_split role counters (a2cu_split_profile)
    uint64_t split_launches = 0;
    std::vector<uint32_t> gstamp;
    // drop-in ("block") mode
    std::vector<BusCmd> buscmds;
    BusCmd *d_buscmds = nullptr;
    size_t buscmds_cap = 0;
    // runs: the commands of one voice in this flush (one CTA of bus_level each)
    struct HostRun { int level; unsigned count; };
    std::vector<HostRun> runs;
    int cur_run = -1;
    uint32_t flush_serial = 1;
    std::vector<BusCmd> sorted_cmds;
    std::vector<BusRun> sorted_runs;
    BusRun *d_runs = nullptr;
    size_t runs_cap = 0;
    // fbdelay delay lines (units/fbdelay.c:187-188): 2 x kFbdSize int32 each
    std::vector<int *> fbd_free;
    int *d_bacc = nullptr;          // [bus][64][2]
    int bacc_cap = 0, nbbus = 0, prev_nbbus = 0;
    int *d_pmstate = nullptr;       // [pm][8]
    int pm_cap = 0, pm_used = 0;
    std::vector<int> pm_free, pm_deferred;
    int32_t *h_xfer = nullptr;      // pinned, 64 x 2
    // generic units (any replaced unit outside a fused leaf voice, BUS_U_* ops)
    struct GUnit { int kind = 0, nin = 0, nout = 0; bool live = false, has_osc = false; OscMirror osc; int *fbd = nullptr; };
    std::vector<GUnit> gunits;
    std::vector<int> gunit_free, gunit_deferred;
    int *d_ustate = nullptr;        // [unit][kUnitWords]
    int ustate_cap = 0;
    // multi-GPU root-bus exchange over peer memory (a2cu_xchg_*, kernels: xchg_root_bus)
    struct Xchg {
        bool enabled = false;
        int world = 1, rank = 0, max_frames = 0;
        unsigned epoch = 0;
        void *base = nullptr;               // own symmetric buffer: 256 B of flags, then data
        void *peer[kMaxPeers] = {nullptr};  // every rank's buffer as mapped here
        bool ipc_opened[kMaxPeers] = {false};
        unsigned *h_status = nullptr;       // mapped pinned word, written by the kernel on timeout
        unsigned long long timeout_cycles = 4000000000ull;
        // lagged mode (a2cu_submit): the window whose root bus is published but whose root stage
        // has not run yet - the next launch (or a2cu_collect) finishes it
        int *d_sum = nullptr;               // [max_frames][2]
        struct Pending {
            bool valid = false;
            unsigned epoch = 0;
            int W = 0, buffer = 0, nsplits = 0;
            int splits[kMaxSplits] = {0};
            int32_t *master = nullptr;
        } pending;
        int deferred_slot = -1;             // result slot (ticket) that waits for `pending`
    } xchg;
};
// Grow a device buffer, keeping the old one if the allocation fails.
template <class T>
static int grow_device(a2cu_engine *e, T **buf, size_t *cap, size_t need_elems, bool sync_first) {
    if (need_elems <= *cap) return A2CU_OK;
    if (sync_first) CK(cudaStreamSynchronize(e->stream));
    T *n = nullptr;
    const size_t ncap = need_elems * 2;
    if (cudaMalloc(&n, ncap * sizeof(T)) != cudaSuccess)
        return fail(A2CU_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(cudaGetLastError()));
    if (*buf) cudaFree(*buf);
    *buf = n;
    *cap = ncap;
    return A2CU_OK;
}

static const size_t kXchgFlagBytes = 256;
static void xchg_release(a2cu_engine *e) {
    a2cu_engine::Xchg &x = e->xchg;
    for (int r = 0; r < kMaxPeers; ++r) {
        if (x.ipc_opened[r] && x.peer[r]) cudaIpcCloseMemHandle(x.peer[r]);
        x.peer[r] = nullptr; x.ipc_opened[r] = false;
    }
    if (x.base) cudaFree(x.base);
    if (x.d_sum) cudaFree(x.d_sum);
    if (x.h_status) cudaFreeHost(x.h_status);
    x = a2cu_engine::Xchg();
}

static XchgParams xchg_params(a2cu_engine *e) {
    XchgParams X;
    memset(&X, 0, sizeof(X));
    a2cu_engine::Xchg &x = e->xchg;
    if (!x.enabled || x.world < 2) return X;
    X.world = x.world; X.rank = x.rank; X.max_frames = x.max_frames;
    X.epoch = ++x.epoch;
    for (int r = 0; r < x.world; ++r) {
        X.flags[r] = (unsigned *)x.peer[r];
        X.data[r] = (int *)((char *)x.peer[r] + kXchgFlagBytes);
    }
    X.status = x.h_status;
    X.timeout_cycles = x.timeout_cycles;
    return X;
}
// Fill the "previous window" half of X from the pending record.
static void xchg_fill_prev(a2cu_engine *e, XchgParams &X) {
    const a2cu_engine::Xchg::Pending &p = e->xchg.pending;
    X.prev_valid = p.valid ? 1 : 0;
    X.prev_epoch = p.epoch; X.prev_W = p.W; X.prev_buffer = p.buffer; X.prev_nsplits = p.nsplits;
    for (int i = 0; i < kMaxSplits; ++i) X.prev_splits[i] = p.splits[i];
    X.prev_master = p.master; X.sum = e->xchg.d_sum;
}
// The pending window's result slot is complete once the work queued so far has run.
extern "C" {
static int xchg_pending_done(a2cu_engine *e);
// Finish the pending window with its own small launch (nothing follows, or what follows cannot).
static int xchg_drain_pending(a2cu_engine *e);
}
static int xchg_check(a2cu_engine *e) {
    if (e->xchg.h_status && *(volatile unsigned *)e->xchg.h_status) {
        char b[64];
        snprintf(b, sizeof(b), "window %u", *(volatile unsigned *)e->xchg.h_status);
        *(volatile unsigned *)e->xchg.h_status = 0;
        return fail(A2CU_ECUDA, "root-bus exchange timed out waiting for a peer (%s)", b);
    }
    return A2CU_OK;
}


void OscMirror::set_phase(const a2cu_engine *e, int phv, unsigned sst) {   // wtosc.c:378-387
    if (wave < 0) { phase = 0; return; }
    phv = h_add(phv, (int)((sst * (dphase >> 8)) >> 8));
    phase = (unsigned long long)(((long long)phv * (long long)e->waves[wave].period) << 8);
}

void OscMirror::write(const a2cu_engine *e, int reg, int v, int start, int dur) {    // WtOsc::write
    switch (reg) {
    case 0:
        wave = v;
        if (v < 0) mode = OSC_OFF;
        else {
            int t = e->waves[v].type;
            mode = t == A2CU_WNOISE ? OSC_NOISE : t == A2CU_WWAVE ? OSC_NOMIP : t == A2CU_WMIPWAVE ? OSC_MIP : OSC_OFF;
            if (mode == OSC_OFF) wave = -1;
        }
        break;
    case 1:
        hramp_set(p, v, start, dur);
        if (!dur) p_ramping = 1;
        break;
    case 3: set_phase(e, v, (unsigned)start); break;
    default: break;     // amplitude does not influence phase or draw count
    }
}

int OscMirror::segment(const a2cu_engine *e, int frames, bool *is_noise) {     // WtOsc::prepare + loop + finish
    *is_noise = false;
    if (mode == OSC_MIP || mode == OSC_NOMIP) {
        if (!e->waves[wave].size[0]) { wave = -1; mode = OSC_OFF; return 0; }
    }
    switch (mode) {
    case OSC_OFF:
        hramp_prepare(p, frames); hramp_run(p, frames);
        return 0;
    case OSC_NOISE: {
        run_pitch(frames);
        int draws = 0;
        for (int i = 0; i < frames; ++i) {      // wtosc.c:140-145
            unsigned long long nph = phase + dphase;
            if ((dphase >= (1u << 23)) || ((nph ^ phase) >> 23)) ++draws;
            phase = nph;
        }
        *is_noise = true;
        return draws;
    }
    case OSC_MIP: {
        const HostWave &w = e->waves[wave];
        run_pitch(frames);
        unsigned est = ((dphase + 255) >> 8) * w.period;
        int m = 0;
        for (; (est > (unsigned)(kMaxPhInc << 8)) && (m < kMipLevels - 1); ++m) est >>= 1;
        unsigned long long ph = phase >> m;
        unsigned dph = (unsigned)(((unsigned long long)dphase * w.period) >> m);
        if (w.flags & A2CU_LOOPED) ph %= (unsigned long long)w.size[m] << 24;
        else if ((ph >> 24) > (unsigned long long)(w.size[m] + kWavePre)) return 0;
        ph += (unsigned long long)dph * (unsigned)frames;      // muted or played: same advance
        phase = ph << m;
        return 0;
    }
    case OSC_NOMIP: {
        const HostWave &w = e->waves[wave];
        run_pitch(frames);
        unsigned long long dp = (unsigned long long)dphase * w.period;
        if (dp >> 32) { phase += dp * (unsigned)frames; return 0; }
        unsigned dph = (unsigned)dp;
        if (dp > (unsigned long long)(kMaxPhInc << 16)) {
            unsigned long long ph = phase;
            for (int i = 0; i < frames; ++i) {
                if (w.flags & A2CU_LOOPED) ph %= (unsigned long long)w.size[0] << 24;
                else if ((ph >> 24) >= w.size[0]) break;
                ph += dph;
            }
            phase = ph;
            return 0;
        }
        if (w.flags & A2CU_LOOPED) {
            unsigned m32 = w.size[0] << 24;
            if (m32) phase %= m32;
        } else if ((phase >> 24) > (unsigned long long)(w.size[0] + kWavePre)) return 0;
        phase += (unsigned long long)dph * (unsigned)frames;
        return 0;
    }
    }
    return 0;
}

static inline void lcg_advance(uint32_t *st, int draws) {     // a2_dsp.h:39-40
    for (int i = 0; i < draws; ++i) *st = *st * 1566083941u + 1u;
}

static int unit_words(const a2cu_unitspec &u) {
    switch (u.kind) {
    case A2CU_WTOSC: return 14;
    case A2CU_PANMIX: return 8;
    case A2CU_FILTER12: return 12 + 2 * u.ninputs;
    case A2CU_WAVESHAPER: return 4;
    case A2CU_FBDELAY: return 10;
    case A2CU_LIMITER: return 3;
    case A2CU_DCBLOCK: return 1 + 2 * u.ninputs;
    case A2CU_DC: return 5;
    default: {
        static const int nops[8] = {1, 2, 3, 4, 3, 4, 2, 4};
        return 16 * nops[u.kind - A2CU_FM1];
    }
    }
}

// Any struct of replaced units (0-2 scratch channels between them) that has no fused kernel runs
// on render_generic. fbdelay is the exception: it is a CTA-cooperative bus unit (a2cu_bus.cuh).
static bool generic_ok(const a2cu_unitspec *c, int n) {
    if (n < 1 || n > kMaxChain) return false;
    for (int i = 0; i < n; ++i) {
        const int k = c[i].kind, ni = c[i].ninputs, no = c[i].noutputs;
        const bool gen = k == A2CU_WTOSC || k == A2CU_DC || (k >= A2CU_FM1 && k <= A2CU_FM4R);
        const bool matchio = k == A2CU_FILTER12 || k == A2CU_WAVESHAPER || k == A2CU_LIMITER || k == A2CU_DCBLOCK;
        if (!gen && !matchio && k != A2CU_PANMIX) return false;
        if (no < 1 || no > 2 || ni < 0 || ni > 2) return false;
        if (gen && (ni != 0 || (k != A2CU_DC && no != 1))) return false;
        if (matchio && (ni != no || ni < 1)) return false;
        if (k == A2CU_PANMIX && ni < 1) return false;
    }
    return true;
}
static bool find_kernel(const a2cu_unitspec *chain, int n, KernelEntry *k, GenericChain *g, bool *generic) {
    auto it = registry().find(sig_of(chain, n));
    if (it != registry().end()) { *k = it->second; *generic = false; return true; }
    if (!generic_ok(chain, n)) return false;
    memset(k, 0, sizeof(*k));
    memset(g, 0, sizeof(*g));
    int w = 0;
    g->n = n;
    for (int i = 0; i < n; ++i) {
        g->kind[i] = chain[i].kind; g->nin[i] = chain[i].ninputs; g->nout[i] = chain[i].noutputs;
        g->add[i] = (chain[i].add ? 1 : 0) | (chain[i].wireout ? 2 : 0);
        g->word[i] = w;
        w += unit_words(chain[i]);
    }
    k->fn = nullptr; k->words = w + 1; k->name = "generic";
    *generic = true;
    return true;
}

// Read the control state of a voice's oscillators back from the device.
static int mirror_create(a2cu_engine *e, int bank, int slot, VoiceMirror **out) {
    Bank *b = e->banks[bank];
    uint64_t key = ((uint64_t)bank << 32) | (uint32_t)slot;
    auto it = e->mirrors.find(key);
    if (it != e->mirrors.end()) { *out = &it->second; return A2CU_OK; }
    VoiceMirror vm;
    CK(cudaStreamSynchronize(e->stream));
    int flags = 0;
    CK(memcpy_done(&flags, b->d_state + slot, sizeof(int), cudaMemcpyDeviceToHost));
    vm.alive = flags & 1;
    int base = 1;
    for (size_t u = 0; u < b->chain.size(); ++u) {
        if (b->chain[u].kind == A2CU_WTOSC) {
            int w[14];
            CK(cudaMemcpy2D(w, sizeof(int), b->d_state + (size_t)base * b->stride + slot, b->stride * sizeof(int),
                            sizeof(int), 14, cudaMemcpyDeviceToHost));
            OscMirror m;
            m.load(w);
            vm.osc.push_back(std::make_pair((int)u, m));
        }
        base += unit_words(b->chain[u]);
    }
    e->mirrors[key] = vm;
    *out = &e->mirrors[key];
    b->set_mirrored(slot, true);
    return A2CU_OK;
}

// Next pinned staging buffer of the ring, at least 'bytes' large. Blocks only if
// the H2D copies issued from that buffer kStageRing windows ago are still
// pending (they never are in steady state).
static int ensure_stage(a2cu_engine *e, size_t bytes) {
    const int k = e->stage_pos;
    e->stage_pos = (k + 1) % a2cu_engine::kStageRing;
    if (!e->stage_done[k]) CK(cudaEventCreateWithFlags(&e->stage_done[k], cudaEventDisableTiming));
    if (e->stage_used[k]) CK(cudaEventSynchronize(e->stage_done[k]));
    if (bytes > e->stage_bufcap[k]) {
        if (e->stage_buf[k]) cudaFreeHost(e->stage_buf[k]);
        e->stage_bufcap[k] = std::max(bytes * 2, (size_t)1 << 20);
        CK(cudaMallocHost(&e->stage_buf[k], e->stage_bufcap[k]));
    }
    e->h_stage = e->stage_buf[k];
    e->stage_cap = e->stage_bufcap[k];
    e->stage_used[k] = true;
    return 0;
}
// Call after the last H2D copy that reads the current staging buffer.
static int stage_release(a2cu_engine *e) {
    const int k = (e->stage_pos + a2cu_engine::kStageRing - 1) % a2cu_engine::kStageRing;
    CK(cudaEventRecord(e->stage_done[k], e->stream));
    return 0;
}

// ---------------------------------------------------------------------------
// Waves (host preparation; semantics of waves.c:59-151 and :629-708)
// ---------------------------------------------------------------------------
static const int kPost = 132;   // A2_WAVEPOST, a2_waves.h:63-64
// Hermite coefficients (16 bytes per sample) only for waves up to this many samples incl. pads;
// longer (sampled) waves are interpolated from the raw int16 data (render_split's raw-tap path)
static const size_t kCoefMaxSamples = 1u << 20;
static const size_t kCoefSlack = 64;    // zero entries after the last table: render_split reads ahead

// Make room for `need` more elements in a device arena (contents are kept, new space is zeroed).
template <class T>
static int arena_reserve(a2cu_engine *e, T **buf, size_t *cap, size_t used, size_t need) {
    if (used + need <= *cap) return A2CU_OK;
    CK(cudaStreamSynchronize(e->stream));
    size_t ncap = std::max((used + need) * 2, (size_t)1 << 16);
    T *n = nullptr;
    if (cudaMalloc(&n, ncap * sizeof(T)) != cudaSuccess)
        return fail(A2CU_ENOMEM, "cudaMalloc wave pool: %s", cudaGetErrorString(cudaGetLastError()));
    CK(memset_done(n, 0, ncap * sizeof(T)));
    if (*buf) {
        CK(memcpy_done(n, *buf, used * sizeof(T), cudaMemcpyDeviceToDevice));
        cudaFree(*buf);
    }
    *buf = n;
    *cap = ncap;
    return A2CU_OK;
}

// Geometry of a new wave (waves.c:59-88): level l holds ceil(length / 2^l) samples plus the pads.
static void wave_layout(a2cu_engine *e, HostWave &w, const unsigned *sizes, int levels) {
    size_t pos = e->pool_used;
    for (int l = 0; l < kMipLevels; ++l) {
        w.size[l] = 0; w.first[l] = 0; w.span[l] = 0; w.coff[l] = -1;
        if (l < levels) {
            w.size[l] = sizes[l];
            w.first[l] = pos;
            w.span[l] = (size_t)kWavePre + sizes[l] + kPost;
            pos += w.span[l];
        }
    }
    w.cbegin = -1; w.ccount = 0;
}

// Pads (unless the host already made them), and the coefficient tables, on the device.
static int wave_finish_on_device(a2cu_engine *e, HostWave &w, int levels, bool build_levels) {
    const int looped = (w.flags & A2CU_LOOPED) ? 1 : 0;
    for (int l = 0; l < levels; ++l) {
        int16_t *lvl = e->d_pool + w.first[l];
        if (build_levels) {
            if (l > 0 && w.size[l]) {       // waves.c:108-151: level l from level l - 1 (already padded)
                const int16_t *src = e->d_pool + w.first[l - 1] + kWavePre;
                wave_mip<<<(w.size[l] + 255) / 256, 256, 0, e->stream>>>(src, lvl + kWavePre, w.size[l]);
                ++e->launches;
            }
            wave_pad<<<1, 256, 0, e->stream>>>(lvl, w.size[l], looped);     // waves.c:90-106
            ++e->launches;
        }
    }
    if (levels && w.total() <= kCoefMaxSamples) {
        size_t need = 0;
        for (int l = 0; l < levels; ++l) need += (size_t)w.size[l] + kPost - 2;
        int r = arena_reserve(e, &e->d_cpool, &e->cpool_cap, e->cpool_used, need + kCoefSlack);
        if (r) return r;
        w.cbegin = (int)e->cpool_used;
        for (int l = 0; l < levels; ++l) {
            const int n = (int)w.size[l] + kPost - 2;
            w.coff[l] = (int)e->cpool_used;
            wave_coef<<<(n + 255) / 256, 256, 0, e->stream>>>(e->d_pool + w.first[l] + kWavePre, e->d_cpool + e->cpool_used, n);
            ++e->launches;
            e->cpool_used += (size_t)n;
        }
        w.ccount = (int)e->cpool_used - w.cbegin;
    }
    CK(cudaGetLastError());
    return A2CU_OK;
}

// The wave descriptor table is the only thing left to (re)upload when waves change.
static int upload_waves(a2cu_engine *e) {
    if (!e->waves_dirty) return 0;
    std::vector<WaveDesc> desc(e->waves.size());
    for (size_t i = 0; i < e->waves.size(); ++i) {
        const HostWave &w = e->waves[i];
        desc[i].type = w.type; desc[i].flags = w.flags; desc[i].period = w.period;
        for (int l = 0; l < kMipLevels; ++l) {
            desc[i].size[l] = w.size[l];
            desc[i].offset[l] = (unsigned)(w.first[l] + kWavePre);
            desc[i].coff[l] = w.coff[l];
        }
    }
    CK(cudaStreamSynchronize(e->stream));
    if (desc.size() > e->waves_cap) {
        if (e->d_waves) cudaFree(e->d_waves);
        e->waves_cap = desc.size() * 2 + 8;
        CK(cudaMalloc(&e->d_waves, e->waves_cap * sizeof(WaveDesc)));
    }
    if (!e->d_pool) {       // kernels take the pointers even when no wave has samples
        int r = arena_reserve(e, &e->d_pool, &e->pool_cap, 0, 16);
        if (r) return r;
    }
    if (!e->d_cpool) {
        int r = arena_reserve(e, &e->d_cpool, &e->cpool_cap, 0, kCoefSlack);
        if (r) return r;
    }
    if (!desc.empty())
        CK(memcpy_done(e->d_waves, desc.data(), desc.size() * sizeof(WaveDesc), cudaMemcpyHostToDevice));
    e->waves_dirty = false;
    return 0;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

const char *a2cu_last_error(void) { return g_err; }

a2cu_engine *a2cu_open(int device, int samplerate, int channels) {
    register_all();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        fail(A2CU_ENODEVICE, "no CUDA device %s (this library has no CPU fallback)", "");
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        fail(A2CU_ENODEVICE, "cudaSetDevice failed%s", "");
        return nullptr;
    }
    register_split();
    a2cu_engine *e = new a2cu_engine();
    e->use_split = getenv("A2CU_NO_SPLIT") == nullptr;
    e->env_stats = getenv("A2CU_STATS") != nullptr;
    e->env_no_copy_stream = getenv("A2CU_NO_COPY_STREAM") != nullptr;
    e->env_no_fuse = getenv("A2CU_NO_FUSE") != nullptr;
    e->env_no_stage = getenv("A2CU_NO_STAGE") != nullptr;
    e->env_one_set = getenv("A2CU_ONE_SET") != nullptr;
    e->env_dry = getenv("A2CU_DRY") != nullptr;
    e->env_no_side_streams = getenv("A2CU_NO_SIDE_STREAMS") != nullptr;
    e->env_split_always = getenv("A2CU_SPLIT_ALWAYS") != nullptr;
    cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e->sm_count <= 0) e->sm_count = 148;
    e->noise_ptr = &e->noiseseed;
    e->device = device;
    e->samplerate = samplerate;
    e->channels = channels < 2 ? 1 : 2;
    // audiality2.c:398-399 (a2_F2Pf: pitch.c:45-48) and :499
    e->basepitch = (int)((float)log2(261.626f / (float)samplerate) * 65536.0f + 0.5f);
    e->msdur = (uint32_t)(samplerate * 65.536f + .5f);
    const HostTables &t = tables();
    const std::vector<int> &f12 = f12_table(samplerate);
    bool ok = cudaMalloc(&e->d_ptab, sizeof(t.ptab)) == cudaSuccess &&
              cudaMalloc(&e->d_f12tab, f12.size() * sizeof(int)) == cudaSuccess &&
              memcpy_done(e->d_f12tab, f12.data(), f12.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMalloc(&e->d_fmsine, sizeof(t.fmsine_packed)) == cudaSuccess &&
              cudaMalloc(&e->d_rstate, 8 * sizeof(int)) == cudaSuccess &&
              memcpy_done(e->d_ptab, t.ptab, sizeof(t.ptab), cudaMemcpyHostToDevice) == cudaSuccess &&
              memcpy_done(e->d_fmsine, t.fmsine_packed, sizeof(t.fmsine_packed), cudaMemcpyHostToDevice) == cudaSuccess;
    // root panmix: vol 1.0, pan 0 (panmix.c:252-262)
    int rs[8] = {65536 << 8, 65536 << 8, 0, 0, 0, 0, 0, 0};
    ok = ok && memcpy_done(e->d_rstate, rs, sizeof(rs), cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaEventCreate(&e->ev0) == cudaSuccess && cudaEventCreate(&e->ev1) == cudaSuccess &&
         cudaEventCreate(&e->ev2) == cudaSuccess;
    if (!ok) {
        fail(A2CU_ECUDA, "a2cu_open: %s", cudaGetErrorString(cudaGetLastError()));
        delete e;
        return nullptr;
    }
    return e;
}

void a2cu_close(a2cu_engine *e) {
    if (!e) return;
    if (e->env_stats && e->ws.windows)
        fprintf(stderr, "a2cu window stats: %ld windows, host us per window: collect/sort %.1f, stage+H2D %.1f, "
                        "render launch %.1f, bus stage launch %.1f\n", e->ws.windows, e->ws.prep_us / e->ws.windows,
                e->ws.stage_us / e->ws.windows, e->ws.api_us / e->ws.windows, e->ws.mix_us / e->ws.windows);
    if (e->env_stats)
        fprintf(stderr, "a2cu stats: %ld flushes %.1f us avg (host side), %ld downloads, wait+copy %.1f us avg, "
                        "%ld launches\n", e->bs.flushes, e->bs.flushes ? e->bs.flush_us / e->bs.flushes : 0.0,
                e->bs.downloads, e->bs.downloads ? e->bs.sync_us / e->bs.downloads : 0.0, (long)e->launches);
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    xchg_release(e);
    for (Bank *b : e->banks) {
        cudaFree(b->d_state); cudaFree(b->d_bus); cudaFree(b->d_noise); cudaFree(b->d_scratch);
        cudaFree(b->d_ev); cudaFree(b->d_runs);
        if (b->ev_consumed) cudaEventDestroy(b->ev_consumed);
        delete b;
    }
    cudaFree(e->d_waves); cudaFree(e->d_pool); cudaFree(e->d_cpool); cudaFree(e->d_ptab); cudaFree(e->d_fmsine); cudaFree(e->d_f12tab);
    cudaFree(e->d_gstate); cudaFree(e->d_rstate); cudaFree(e->d_mixev);
    cudaFree(e->d_acc); cudaFree(e->d_master);
    cudaFree(e->d_buscmds); cudaFree(e->d_bacc); cudaFree(e->d_pmstate); cudaFree(e->d_ustate); cudaFree(e->d_runs); cudaFree(e->d_fuse_counter);
    for (auto &g : e->gunits) if (g.fbd) cudaFree(g.fbd);
    for (int *p : e->fbd_free) cudaFree(p);
    if (e->h_xfer) cudaFreeHost(e->h_xfer);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    for (int k = 0; k < a2cu_engine::kSideStreams; ++k) {
        if (e->side[k]) cudaStreamDestroy(e->side[k]);
        if (e->side_join[k]) cudaEventDestroy(e->side_join[k]);
    }
    if (e->side_fork) cudaEventDestroy(e->side_fork);
    for (int k = 0; k < a2cu_engine::kStageRing; ++k) {
        if (e->stage_buf[k]) cudaFreeHost(e->stage_buf[k]);
        if (e->stage_done[k]) cudaEventDestroy(e->stage_done[k]);
    }
    for (auto &sl : e->slots) {
        if (sl.h_out) cudaFreeHost(sl.h_out);
        if (sl.ev0) cudaEventDestroy(sl.ev0);
        if (sl.ev1) cudaEventDestroy(sl.ev1);
        if (sl.ev2) cudaEventDestroy(sl.ev2);
        if (sl.done) cudaEventDestroy(sl.done);
    }
    if (e->h_out) cudaFreeHost(e->h_out);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->ev2) cudaEventDestroy(e->ev2);
    delete e;
}

int a2cu_set_stream(a2cu_engine *e, void *s) {
    if (!e) return A2CU_EINVAL;
    e->stream = (cudaStream_t)s;
    return A2CU_OK;
}
int a2cu_basepitch(const a2cu_engine *e) { return e->basepitch; }
uint32_t a2cu_msdur(const a2cu_engine *e) { return e->msdur; }
uint64_t a2cu_now(const a2cu_engine *e) { return e->now; }
int a2cu_set_root_wake_period(a2cu_engine *e, uint32_t p) { e->root_wake = p; return A2CU_OK; }
int a2cu_set_noiseseed(a2cu_engine *e, uint32_t s) { *e->noise_ptr = s; return A2CU_OK; }
int a2cu_set_noise_state_ptr(a2cu_engine *e, uint32_t *p) {
    if (!e) return A2CU_EINVAL;
    e->noise_ptr = p ? p : &e->noiseseed;
    return A2CU_OK;
}
uint64_t a2cu_launch_count(const a2cu_engine *e) { return e->launches; }
uint64_t a2cu_split_launch_count(const a2cu_engine *e) { return e->split_launches; }
int a2cu_set_split(a2cu_engine *e, int on) { e->use_split = on != 0; return A2CU_OK; }
static const size_t kProfWords = 8 + 6 * 64 * 2;    // role counters + timeline of CTA 0 (a2cu_split.cuh)
int a2cu_split_profile(a2cu_engine *e, int enable, uint64_t out[8]) {
    if (!e) return A2CU_EINVAL;
    cudaSetDevice(e->device);
    if (out && e->d_prof) {
        CK(cudaStreamSynchronize(e->stream));
        CK(memcpy_done(out, e->d_prof, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    }
    if (enable && !e->d_prof) {
        CK(cudaMalloc(&e->d_prof, kProfWords * sizeof(uint64_t)));
    }
    if (e->d_prof) CK(memset_done(e->d_prof, 0, 8 * sizeof(uint64_t)));
    if (!enable && e->d_prof) { cudaFree(e->d_prof); e->d_prof = nullptr; }
    return A2CU_OK;
}
// Reset the grid-wide wall-clock marks of the timeline (min / max slots) before a launch.
int a2cu_split_trace_reset(a2cu_engine *e) {
    if (!e || !e->d_prof) return A2CU_EINVAL;
    cudaSetDevice(e->device);
    uint64_t init[8] = {~0ull, 0, 0, 0, 0, 0, 0, 0};
    CK(cudaMemcpyAsync(e->d_prof + 8 + (4 * 64 + 56) * 2, init, sizeof(init), cudaMemcpyHostToDevice, e->stream));
    return A2CU_OK;
}
// Timeline of CTA 0 / voice set 0 of the LAST render_split launch while profiling is armed:
// out[((role * 64 + fragment) * 2 + end)] cycles since the pipeline start; roles 0 control,
// 1 filter recurrence, 2 / 3 stage A / C of helper 0, 4 / 5 of the last helper. 768 words.
int a2cu_split_trace(a2cu_engine *e, uint64_t *out) {
    if (!e || !out || !e->d_prof) return A2CU_EINVAL;
    cudaSetDevice(e->device);
    CK(cudaStreamSynchronize(e->stream));
    CK(memcpy_done(out, e->d_prof + 8, (kProfWords - 8) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return A2CU_OK;
}
uint64_t a2cu_h2d_bytes(const a2cu_engine *e) { return e->h2d_bytes; }
uint64_t a2cu_d2h_bytes(const a2cu_engine *e) { return e->d2h_bytes; }
int a2cu_set_timing(a2cu_engine *e, int on) { e->timing = on != 0; return A2CU_OK; }
float a2cu_last_render_ms(a2cu_engine *e) { return e->last_ms; }
float a2cu_last_mix_ms(a2cu_engine *e) { return e->last_mix_ms; }
int a2cu_set_post_root_stage(a2cu_engine *e, int on) { e->post_root = on != 0; return A2CU_OK; }
int a2cu_set_output_format(a2cu_engine *e, int fmt) {
    if (!e || fmt < 0 || fmt > 2) return fail(A2CU_EINVAL, "a2cu_set_output_format: 0 int32, 1 float32, 2 int16%s");
    e->out_fmt = fmt;
    return A2CU_OK;
}
// bytes of one output sample (the raw root bus of the multi-GPU cut is always int32)
static size_t out_sample_bytes(const a2cu_engine *e) { return (e->post_root && e->out_fmt == 2) ? 2 : 4; }

// ---- waves -----------------------------------------------------------------
int a2cu_wave_upload(a2cu_engine *e, int type, unsigned period, unsigned flags, const int16_t *data,
                     unsigned length) {
    if (!e || type < A2CU_WOFF || type > A2CU_WMIPWAVE) return fail(A2CU_EINVAL, "bad wave type%s");
    if ((type == A2CU_WWAVE || type == A2CU_WMIPWAVE) && !data && length) return fail(A2CU_EINVAL, "no data%s");
    cudaSetDevice(e->device);
    HostWave w;
    w.type = type; w.flags = flags; w.period = period;
    const int levels = type == A2CU_WMIPWAVE ? kMipLevels : (type == A2CU_WWAVE ? 1 : 0);
    unsigned sizes[kMipLevels];
    for (int l = 0; l < kMipLevels; ++l) sizes[l] = (length + (1u << l) - 1) >> l;
    size_t need = 0;
    for (int l = 0; l < levels; ++l) need += (size_t)kWavePre + sizes[l] + kPost;
    int r = arena_reserve(e, &e->d_pool, &e->pool_cap, e->pool_used, need + 16);
    if (r) return r;
    wave_layout(e, w, sizes, levels);
    if (levels) {
        // only the raw samples cross the bus; pads, mip levels and coefficients are made on the device
        if (length)
            CK(cudaMemcpyAsync(e->d_pool + w.first[0] + kWavePre, data, (size_t)length * sizeof(int16_t),
                               cudaMemcpyHostToDevice, e->stream));
        e->h2d_bytes += (size_t)length * sizeof(int16_t);
        e->pool_used += need;
        r = wave_finish_on_device(e, w, levels, true);
        if (r) return r;
    }
    e->waves.push_back(std::move(w));
    e->waves_dirty = true;
    return (int)e->waves.size() - 1;
}

int a2cu_wave_upload_prepared(a2cu_engine *e, int type, unsigned period, unsigned flags,
                              const int16_t *const *data, const unsigned *size) {
    if (!e) return A2CU_EINVAL;
    cudaSetDevice(e->device);
    HostWave w;
    w.type = type; w.flags = flags; w.period = period;
    const int levels = type == A2CU_WMIPWAVE ? kMipLevels : (type == A2CU_WWAVE ? 1 : 0);
    size_t need = 0;
    for (int l = 0; l < levels; ++l) need += (size_t)kWavePre + size[l] + kPost;
    int r = arena_reserve(e, &e->d_pool, &e->pool_cap, e->pool_used, need + 16);
    if (r) return r;
    wave_layout(e, w, size, levels);
    if (levels) {
        // the host's own prepared buffers, pads included, exactly as they are (a2_waves.h:88-103)
        for (int l = 0; l < levels; ++l) {
            CK(cudaMemcpyAsync(e->d_pool + w.first[l], data[l], w.span[l] * sizeof(int16_t), cudaMemcpyHostToDevice,
                               e->stream));
            e->h2d_bytes += w.span[l] * sizeof(int16_t);
        }
        CK(cudaStreamSynchronize(e->stream));       // the host may free or rewrite its buffers
        e->pool_used += need;
        r = wave_finish_on_device(e, w, levels, false);
        if (r) return r;
    }
    e->waves.push_back(std::move(w));
    e->waves_dirty = true;
    return (int)e->waves.size() - 1;
}

int a2cu_wave_builtin(a2cu_engine *e, const char *name) {
    if (!e || !name) return A2CU_EINVAL;
    for (size_t i = 0; i < e->waves.size(); ++i)
        if (e->waves[i].name == name) return (int)i;
    const int N = 2048;                     // A2_WAVEPERIOD, a2_waves.h:71
    std::vector<int16_t> buf(N, 0);
    int h;
    std::string n(name);
    if (n == "off") h = a2cu_wave_upload(e, A2CU_WOFF, 0, 0, nullptr, 0);
    else if (n == "noise") h = a2cu_wave_upload(e, A2CU_WNOISE, 256, A2CU_LOOPED, nullptr, 0);
    else {
        int duty = n == "square" ? 50 : (n.compare(0, 5, "pulse") == 0 ? atoi(name + 5) : 0);
        if (duty > 0 && duty <= 50) {
            // waves.c:637-651; index s1 keeps the previous duty cycle's -32767
            int s1 = (N * duty + 50) / 100;
            for (int s = 0; s < N; ++s) buf[s] = s < s1 ? 32767 : -32767;
        } else if (n == "saw") {
            for (int s = 0; s < N; ++s) buf[s] = (int16_t)(s * 65534 / N - 32767);
        } else if (n == "triangle") {
            for (int s = 0; s < N; ++s) buf[s] = (int16_t)(s * 65534 / N - 32767);
            for (int s = 0; s < N / 2; ++s)
                buf[(5 * N / 4 - s - 1) % N] = buf[s + N / 4] = (int16_t)(s * 65534 * 2 / N - 32767);
        } else if (n == "sine" || n == "asine" || n == "hsine" || n == "qsine") {
            for (int s = 0; s < N; ++s) buf[s] = (int16_t)(sin(s * 2.0f * M_PI / N) * 32767.0f);
            if (n != "sine")
                for (int s = N / 2; s < N; ++s) buf[s] = (int16_t)-buf[s];
            if (n == "hsine" || n == "qsine")
                for (int s = N / 2; s < N; ++s) buf[s] = 0;
            if (n == "qsine")
                for (int s = 0; s < N / 4; ++s) buf[s + N / 2] = buf[s];
        } else
            return fail(A2CU_EINVAL, "unknown builtin wave '%s'", name);
        h = a2cu_wave_upload(e, A2CU_WMIPWAVE, N, A2CU_LOOPED, buf.data(), N);
    }
    if (h >= 0) e->waves[h].name = n;
    return h;
}

int a2cu_wave_unload(a2cu_engine *e, int wave) {
    if (!e || wave < 0 || wave >= (int)e->waves.size()) return A2CU_EINVAL;
    e->waves[wave].size[0] = 0;
    e->waves_dirty = true;
    return A2CU_OK;
}

int a2cu_wave_read(a2cu_engine *e, int wave, int level, int16_t *out, unsigned cap, unsigned *size) {
    if (!e || wave < 0 || wave >= (int)e->waves.size() || level < 0 || level >= kMipLevels) return A2CU_EINVAL;
    const HostWave &w = e->waves[wave];
    if (size) *size = w.size[level];
    unsigned n = (unsigned)std::min((size_t)cap, w.span[level]);
    if (out && n) {
        cudaSetDevice(e->device);
        CK(cudaStreamSynchronize(e->stream));
        CK(memcpy_done(out, e->d_pool + w.first[level], (size_t)n * sizeof(int16_t), cudaMemcpyDeviceToHost));
    }
    return (int)(out ? n : w.span[level]);
}

// ---- groups and banks ------------------------------------------------------
int a2cu_group_new(a2cu_engine *e) {
    if (!e) return A2CU_EINVAL;
    cudaSetDevice(e->device);
    if (e->ngroups + 1 > e->gstate_cap) {
        int ncap = std::max(64, e->gstate_cap * 2);
        int *n = nullptr;
        CK(cudaMalloc(&n, (size_t)ncap * 8 * sizeof(int)));
        CK(cudaStreamSynchronize(e->stream));
        if (e->d_gstate) {
            CK(memcpy_done(n, e->d_gstate, (size_t)e->ngroups * 8 * sizeof(int), cudaMemcpyDeviceToDevice));
            cudaFree(e->d_gstate);
        }
        e->d_gstate = n;
        e->gstate_cap = ncap;
    }
    int gs[8] = {65536 << 8, 65536 << 8, 0, 0, 0, 0, 0, 0};
    CK(memcpy_done(e->d_gstate + (size_t)e->ngroups * 8, gs, sizeof(gs), cudaMemcpyHostToDevice));
    e->gstamp.push_back(e->stamp++);
    return e->ngroups++;
}

int a2cu_chain_supported(const a2cu_unitspec *chain, int nunits) {
    register_all();
    return (registry().count(sig_of(chain, nunits)) || generic_ok(chain, nunits)) ? 1 : 0;
}

// Argument of a unit's EV_INIT / BUS_U_INIT record: what its Initialize() reads besides the registers.
static int init_arg(const a2cu_engine *e, int kind, int transpose) {
    if (kind == A2CU_WTOSC || kind >= A2CU_FM1) return transpose + e->basepitch;
    if (kind == A2CU_FILTER12) return transpose;
    if (kind == A2CU_LIMITER) return ((64 << 16) << 8) / e->samplerate;     // limiter.c:165-168, cooked
    return 0;
}

static void push_event(Bank *b, uint64_t time, int voice, int kind, int unit, int reg, int value, uint32_t dur) {
    HostEvent ev;
    ev.time = time; ev.seq = b->seq++; ev.voice = voice;
    ev.y = (uint32_t)kind | ((uint32_t)unit << 8) | ((uint32_t)(reg & 0xff) << 16);
    ev.value = value; ev.dur = dur;
    b->events.push_back(ev);
}

int a2cu_bank_new(a2cu_engine *e, const a2cu_unitspec *chain, int nunits, int nvoices,
                  const int32_t *transpose, const int32_t *group, unsigned substart) {
    if (!e || !chain || nunits < 1 || nvoices < 1) return fail(A2CU_EINVAL, "a2cu_bank_new: bad args%s");
    KernelEntry ke;
    GenericChain gc;
    bool generic = false;
    if (!find_kernel(chain, nunits, &ke, &gc, &generic))
        return fail(A2CU_ENOTIMPL, "no kernel for voice structure %s", sig_of(chain, nunits).c_str());
    cudaSetDevice(e->device);
    Bank *b = new Bank();
    b->chain.assign(chain, chain + nunits);
    b->k = ke;
    b->generic = generic;
    b->gchain = gc;
    b->nvoices = nvoices;
    b->stamp = e->stamp++;
    b->stride = ((size_t)nvoices + kThreads - 1) / kThreads * kThreads;
    b->transpose.assign(nvoices, 0);
    b->group.assign(nvoices, -1);
    if (transpose) b->transpose.assign(transpose, transpose + nvoices);
    if (group) b->group.assign(group, group + nvoices);
    std::vector<int> bus(b->stride, 0);
    for (int i = 0; i < nvoices; ++i) {
        if (b->group[i] >= e->ngroups) { delete b; return fail(A2CU_EINVAL, "bad group index%s"); }
        bus[i] = b->group[i] < 0 ? 0 : 1 + b->group[i];
    }
    for (int u = 0; u < nunits; ++u)
        if (chain[u].kind == A2CU_WTOSC) b->has_noise = true;   // may select the noise wave later
    size_t sbytes = (size_t)b->k.words * b->stride * sizeof(int);
    if (cudaMalloc(&b->d_state, sbytes) != cudaSuccess || cudaMalloc(&b->d_bus, b->stride * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&b->d_noise, b->stride * sizeof(unsigned)) != cudaSuccess) {
        delete b;
        return fail(A2CU_ENOMEM, "cudaMalloc bank state: %s", cudaGetErrorString(cudaGetLastError()));
    }
    CK(memset_done(b->d_state, 0, sbytes));
    CK(memset_done(b->d_noise, 0, b->stride * sizeof(unsigned)));
    CK(memcpy_done(b->d_bus, bus.data(), b->stride * sizeof(int), cudaMemcpyHostToDevice));
    if (b->generic) {
        CK(cudaMalloc(&b->d_scratch, b->stride * kMaxFrag * 2 * sizeof(int)));
        CK(memset_done(b->d_scratch, 0, b->stride * kMaxFrag * 2 * sizeof(int)));
    }
    // Initialize() of every unit, in chain order, at the current time
    uint64_t t = (e->now & ~(uint64_t)0xff) | (substart & 0xff);
    const HostTables &tb = tables();
    for (int i = 0; i < nvoices; ++i) {
        push_event(b, t, i, EV_START, 0, 0, 0, 0);
        for (int u = 0; u < nunits; ++u) {
            int kind = chain[u].kind;
            int arg = init_arg(e, kind, b->transpose[i]);
            push_event(b, t, i, EV_INIT, u, 0, arg, 0);
            if (kind == A2CU_FILTER12)      // exact coefficient from the host libm
                push_event(b, t, i, EV_WRITE, u, 5,
                           tb.f12_coeff((int)((unsigned)arg << 8), e->samplerate), 0);
            if (kind == A2CU_DCBLOCK)       // dcblock.c:127-128: default cutoff -5, cooked
                push_event(b, t, i, EV_WRITE, u, 0,
                           tb.f12_coeff((int)((unsigned)(b->transpose[i] - (5 << 16)) << 8), e->samplerate), 0);
        }
    }
    e->banks.push_back(b);
    return (int)e->banks.size() - 1;
}

static Bank *get_bank(a2cu_engine *e, int bank) {
    if (!e || bank < 0 || bank >= (int)e->banks.size()) return nullptr;
    return e->banks[bank];
}

const char *a2cu_bank_kernel_name(a2cu_engine *e, int bank) {
    Bank *b = get_bank(e, bank);
    return b ? b->k.name : "";
}
int a2cu_bank_state_bytes(a2cu_engine *e, int bank) {
    Bank *b = get_bank(e, bank);
    return b ? b->k.words * (int)sizeof(int) : A2CU_EINVAL;
}

// Translate a VM-level register write into the device-level record(s): the
// host half of each unit's A2_write_cb (see a2cu_device.cuh write()). Returns
// the number of (reg, value, dur) records produced (1 or 2) or a negative error.
struct Cooked { int reg, value; uint32_t dur; };
static int cook(a2cu_engine *e, int kind, int reg, int value, int start, uint32_t dur, int transpose,
                Cooked out[2]) {
    if (reg < 0) return fail(A2CU_EINVAL, "bad register%s");
    int n = 1;
    switch (kind) {
    case A2CU_WTOSC:
        if (reg > 3) return fail(A2CU_EINVAL, "wtosc has 4 registers%s");
        if (reg == 0) {                 // wtosc.c:433-483
            int h = value >> 16;
            int w = (h >= 0 && h < (int)e->waves.size()) ? h : -1;
            if (w >= 0) {
                const HostWave &hw = e->waves[w];
                if ((hw.type == A2CU_WWAVE || hw.type == A2CU_WMIPWAVE) && hw.size[0] > 0x01000000u - 1 - kPost)
                    w = -1;
                if (hw.type == A2CU_WOFF) w = -1;
            }
            value = w;
        } else if (reg == 1)            // wtosc.c:486-492
            value = value + transpose + e->basepitch;
        break;
    case A2CU_PANMIX:
        if (reg > 1) return fail(A2CU_EINVAL, "panmix has 2 registers%s");
        break;
    case A2CU_WAVESHAPER:
        if (reg > 0) return fail(A2CU_EINVAL, "waveshaper has 1 register%s");
        break;
    case A2CU_FBDELAY:              // fbdelay.c:229-270: ms -> frames on the host, gains as they are
        if (reg > 6) return fail(A2CU_EINVAL, "fbdelay has 7 registers%s");
        if (reg < 3) value = (int)((int64_t)value * e->samplerate / 65536000);
        break;
    case A2CU_LIMITER:              // limiter.c:187-199
        if (reg > 1) return fail(A2CU_EINVAL, "limiter has 2 registers%s");
        if (reg == 0) value = (int)((unsigned)value << 8) / e->samplerate;
        else {
            unsigned t = (unsigned)value << 8;
            value = (int)(t < 256 ? 256u : t);
        }
        break;
    case A2CU_DCBLOCK: {            // dcblock.c:109-114: a2_P2I of the 16:16 pitch itself
        if (reg > 0) return fail(A2CU_EINVAL, "dcblock has 1 register%s");
        const int pitch = value + transpose;
        value = tables().f12_coeff((int)((unsigned)pitch << 8), e->samplerate);
        break;
    }
    case A2CU_DC:                   // dc.c:183-247: both registers go to the device as they are
        if (reg > 1) return fail(A2CU_EINVAL, "dc has 2 registers%s");
        break;
    case A2CU_FILTER12:
        if (reg > 4) return fail(A2CU_EINVAL, "filter12 has 5 registers%s");
        if (reg == 0) {                 // filter12.c:141-147
            value = value + transpose;
            if ((uint64_t)dur + (uint64_t)start < 256) {
                // ramper snaps to the target: the coefficient is known here,
                // computed with the host libm exactly like the reference
                out[1].reg = 5;
                out[1].value = tables().f12_coeff((int)((unsigned)value << 8), e->samplerate);
                out[1].dur = 0;
                n = 2;
            }
        } else if (reg == 1)            // filter12.c:149-162
            value = value < 512 ? 32768 : (65536 << 8) / value;
        else
            value >>= 8;                // filter12.c:164-177
        break;
    default: {
        if (kind < A2CU_FM1 || kind > A2CU_FM4R) return fail(A2CU_EINVAL, "unknown unit kind%s");
        static const int nops[8] = {1, 2, 3, 4, 3, 4, 2, 4};
        if (reg > 3 * nops[kind - A2CU_FM1]) return fail(A2CU_EINVAL, "fm register out of range%s");
        if (reg == 1) value = value + transpose + e->basepitch;   // fm.c:417-423
        break;
    }
    }
    out[0].reg = reg; out[0].value = value; out[0].dur = dur;
    return n;
}

static int cook_write(a2cu_engine *e, Bank *b, int voice, int unit, int reg, int value, uint64_t when, uint32_t dur) {
    if (unit < 0 || unit >= (int)b->chain.size()) return fail(A2CU_EINVAL, "bad unit%s");
    Cooked c[2];
    int n = cook(e, b->chain[unit].kind, reg, value, (int)(when & 0xff), dur, b->transpose[voice], c);
    if (n < 0) return n;
    if (b->chain[unit].kind == A2CU_WTOSC && c[0].reg == 0 && c[0].value >= 0) {
        const HostWave &hw = e->waves[c[0].value];
        const size_t total = hw.total();
        // render_split evaluates oscillators in closed form: fine for every looped or mip-mapped
        // wave (coefficient table, or raw taps from the pool for large sampled waves); the shared
        // noise LCG and the per-sample end check of a one-shot wave above A2_MAXPHINC
        // (wtosc.c:301-358) need the frame-serial kernel
        if (hw.type == A2CU_WNOISE || (hw.type == A2CU_WWAVE && !(hw.flags & A2CU_LOOPED))) b->exotic = true;
        if (hw.type == A2CU_WNOISE) e->noise_seen = true;
        // no Hermite-coefficient table (upload_waves), or a looped plain wave whose per-sample wrapped
        // loop above A2_MAXPHINC reads raw taps (wtosc.c:301-358)
        if (total > ((size_t)1 << 20) || hw.type == A2CU_WWAVE) b->raw_taps = true;
        if (hw.type == A2CU_WMIPWAVE && total <= ((size_t)1 << 20)) b->stage_wave = c[0].value;
    }
    for (int i = 0; i < n; ++i) push_event(b, when, voice, EV_WRITE, unit, c[i].reg, c[i].value, c[i].dur);
    return A2CU_OK;
}

int a2cu_bank_write(a2cu_engine *e, int bank, int voice, int unit, int reg, int32_t value, uint64_t when,
                    uint32_t dur) {
    Bank *b = get_bank(e, bank);
    if (!b || voice < 0 || voice >= b->nvoices) return fail(A2CU_EINVAL, "bad bank/voice%s");
    if (when < e->now) return fail(A2CU_ELATE, "event time already rendered%s");
    return cook_write(e, b, voice, unit, reg, value, when, dur);
}

int a2cu_bank_write_all(a2cu_engine *e, int bank, int unit, int reg, const int32_t *values, int stride,
                        uint64_t when, uint32_t dur) {
    Bank *b = get_bank(e, bank);
    if (!b || !values) return fail(A2CU_EINVAL, "bad bank%s");
    if (when < e->now) return fail(A2CU_ELATE, "event time already rendered%s");
    if (unit < 0 || unit >= (int)b->chain.size()) return fail(A2CU_EINVAL, "bad unit%s");
    b->events.reserve(b->events.size() + (size_t)b->nvoices);
    const int kind = b->chain[unit].kind;
    // registers whose cooked value depends on the voice (transpose) or needs the wave checks
    const bool per_voice = (kind == A2CU_WTOSC && reg <= 1) || (kind == A2CU_FILTER12 && reg == 0) ||
                           (kind >= A2CU_FM1 && reg == 1);
    if (per_voice) {
        for (int i = 0; i < b->nvoices; ++i) {
            int r = cook_write(e, b, i, unit, reg, values[(size_t)i * stride], when, dur);
            if (r) return r;
        }
        return A2CU_OK;
    }
    // bulk path: the cooked record is a function of the value alone
    Cooked c[2];
    int n = cook(e, kind, reg, values[0], (int)(when & 0xff), dur, 0, c);
    if (n < 0) return n;
    Bank::BulkEvent be;
    be.time = when; be.seq = b->seq; b->seq += (uint32_t)b->nvoices;
    be.y = (uint32_t)EV_WRITE | ((uint32_t)unit << 8) | ((uint32_t)(c[0].reg & 0xff) << 16);
    be.dur = c[0].dur; be.scalar = c[0].value; be.has_values = stride != 0;
    if (stride) {
        be.values.resize((size_t)b->nvoices);
        be.values[0] = c[0].value;
        for (int i = 1; i < b->nvoices; ++i) {
            n = cook(e, kind, reg, values[(size_t)i * stride], (int)(when & 0xff), dur, 0, c);
            if (n < 0) return n;
            be.values[i] = c[0].value;
        }
    }
    b->bulk.push_back(std::move(be));
    return A2CU_OK;
}

// Pause / resume a bank: a disabled bank is not rendered (its voices keep their
// state, its pending writes apply when it runs again). Lets one engine serve
// many voice banks round-robin.
int a2cu_bank_enable(a2cu_engine *e, int bank, int enabled) {
    Bank *b = get_bank(e, bank);
    if (!b || b->dynamic) return fail(A2CU_EINVAL, "bad bank%s");
    b->enabled = enabled != 0;
    return A2CU_OK;
}

int a2cu_bank_wake(a2cu_engine *e, int bank, int voice, uint64_t when) {
    Bank *b = get_bank(e, bank);
    if (!b || voice >= b->nvoices) return fail(A2CU_EINVAL, "bad bank/voice%s");
    if (when < e->now) return fail(A2CU_ELATE, "event time already rendered%s");
    if (voice >= 0) push_event(b, when, voice, EV_WAKE, 0, 0, 0, 0);
    else
        for (int i = 0; i < b->nvoices; ++i) push_event(b, when, i, EV_WAKE, 0, 0, 0, 0);
    return A2CU_OK;
}

int a2cu_bank_kill(a2cu_engine *e, int bank, int voice, uint64_t when) {
    Bank *b = get_bank(e, bank);
    if (!b || voice < 0 || voice >= b->nvoices) return fail(A2CU_EINVAL, "bad bank/voice%s");
    if (when < e->now) return fail(A2CU_ELATE, "event time already rendered%s");
    push_event(b, when, voice, EV_STOP, 0, 0, 0, 0);
    return A2CU_OK;
}

static int mix_write(a2cu_engine *e, int target, int reg, int value, uint64_t when, uint32_t dur) {
    if (when < e->now) return fail(A2CU_ELATE, "event time already rendered%s");
    if (reg > 1) return fail(A2CU_EINVAL, "panmix has 2 registers%s");
    MixHostEvent m;
    m.time = when; m.seq = e->mixseq++; m.target = target; m.reg = reg < 0 ? -1 : reg;
    m.value = value; m.dur = dur;
    e->mixev.push_back(m);
    if (target < 0 && reg >= 0) {
        e->root_written = true;
        e->root_until = std::max(e->root_until, when + dur + 512 + ((uint64_t)kMaxFrag << 8));
    }
    if (target >= 0) {
        // the group's wake-up cuts its children's segments (core.c:1769-1776)
        for (Bank *b : e->banks)
            for (int i = 0; i < b->nvoices; ++i)
                if (b->group[i] == target) push_event(b, when, i, EV_WAKE, 0, 0, 0, 0);
    }
    return A2CU_OK;
}
int a2cu_group_write(a2cu_engine *e, int group, int reg, int32_t value, uint64_t when, uint32_t dur) {
    if (!e || group < 0 || group >= e->ngroups) return fail(A2CU_EINVAL, "bad group%s");
    return mix_write(e, group, reg, value, when, dur);
}
int a2cu_root_write(a2cu_engine *e, int reg, int32_t value, uint64_t when, uint32_t dur) {
    if (!e) return A2CU_EINVAL;
    return mix_write(e, -1, reg, value, when, dur);
}

// ---- rendering ---------------------------------------------------------------
static int collect_splits(a2cu_engine *e, uint64_t t0, uint64_t t1, int *splits, int *n) {
    // Root-level segment cuts inside [t0, t1): periodic root wake-ups and root
    // writes. They apply to every voice (the root's inline recursion).
    std::vector<int> s;
    if (e->root_wake) {
        uint64_t k = t0 / e->root_wake;
        for (uint64_t t = k * e->root_wake; t < t1; t += e->root_wake)
            if (t > t0 && t < t1) s.push_back((int)((t - t0) >> 8));
    }
    for (auto &m : e->mixev)
        if (m.target < 0 && m.time > t0 && m.time < t1) s.push_back((int)((m.time - t0) >> 8));
    std::sort(s.begin(), s.end());
    s.erase(std::unique(s.begin(), s.end()), s.end());
    s.erase(std::remove(s.begin(), s.end(), 0), s.end());
    if ((int)s.size() > kMaxSplits) return -1;
    *n = (int)s.size();
    for (int i = 0; i < *n; ++i) splits[i] = s[i];
    return 0;
}

// Bank-mode noise planner: seeds for every noise segment of the window, in
// tree-walk order (newest bank / highest voice index first, like the
// reference's head insertion in a2_VoiceNew, core.c:476-477).
static int plan_noise(a2cu_engine *e, uint64_t t0, int W, int buffer, const int *splits, int nsplits,
                      std::vector<std::vector<HostEvent>> &due) {
    // 1. new mirrors for voices that select the noise wave in this window
    for (size_t bi = 0; bi < e->banks.size(); ++bi) {
        Bank *b = e->banks[bi];
        if (b->dynamic) continue;
        for (const HostEvent &ev : due[bi]) {
            int kind = ev.y & 0xff, unit = (ev.y >> 8) & 0xff, reg = (ev.y >> 16) & 0xff;
            if (kind != EV_WRITE || b->chain[unit].kind != A2CU_WTOSC || reg != 0) continue;
            if (ev.value < 0 || e->waves[ev.value].type != A2CU_WNOISE) continue;
            VoiceMirror *vm;
            int r = mirror_create(e, (int)bi, ev.voice, &vm);
            if (r) return r;
        }
    }
    if (e->mirrors.empty()) return A2CU_OK;
    // 2. mirrored voices in walk order
    struct Item { uint32_t rootkey, bstamp; int bank, slot; VoiceMirror *vm; size_t lo, hi; };
    std::vector<Item> items;
    for (auto &kv : e->mirrors) {
        Item it;
        it.bank = (int)(kv.first >> 32); it.slot = (int)(uint32_t)kv.first; it.vm = &kv.second;
        Bank *b = e->banks[it.bank];
        if (b->dynamic) continue;
        int g = b->group[it.slot];
        it.bstamp = b->stamp;
        it.rootkey = g >= 0 ? e->gstamp[g] : b->stamp;
        auto cmp = [](const HostEvent &a, int v) { return a.voice < v; };
        it.lo = std::lower_bound(due[it.bank].begin(), due[it.bank].end(), it.slot, cmp) - due[it.bank].begin();
        it.hi = std::lower_bound(due[it.bank].begin(), due[it.bank].end(), it.slot + 1, cmp) - due[it.bank].begin();
        it.vm->cursor = it.lo;
        items.push_back(it);
    }
    std::sort(items.begin(), items.end(), [](const Item &a, const Item &c) {
        if (a.rootkey != c.rootkey) return a.rootkey > c.rootkey;
        if (a.bstamp != c.bstamp) return a.bstamp > c.bstamp;
        return a.slot > c.slot;
    });
    std::vector<std::vector<HostEvent>> seeds(e->banks.size());
    // 3. pieces = fragments cut by the root-level splits; voices inside a piece
    for (int a = 0; a < W;) {
        int pos = a % buffer;
        int pe = a - pos + std::min(buffer, (pos / kMaxFrag + 1) * kMaxFrag);
        pe = std::min(pe, W);
        for (int k = 0; k < nsplits; ++k)
            if (splits[k] > a) pe = std::min(pe, splits[k]);
        for (Item &it : items) {
            Bank *b = e->banks[it.bank];
            std::vector<HostEvent> &ev = due[it.bank];
            VoiceMirror *vm = it.vm;
            int sfr = a;
            while (sfr < pe) {
                while (vm->cursor < it.hi && (int)((ev[vm->cursor].time - t0) >> 8) <= sfr) {
                    const HostEvent &x = ev[vm->cursor++];
                    int kind = x.y & 0xff, unit = (x.y >> 8) & 0xff, reg = (x.y >> 16) & 0xff;
                    int start = (int)(x.time & 0xff);
                    if (kind == EV_START) vm->alive = 1;
                    else if (kind == EV_STOP) vm->alive = 0;
                    else if (kind == EV_INIT || kind == EV_WRITE)
                        for (auto &om : vm->osc)
                            if (om.first == unit) {
                                if (kind == EV_INIT) om.second.init(e, x.value, (unsigned)start);
                                else om.second.write(e, reg, x.value, start, (int)x.dur);
                            }
                }
                int nxt = pe;
                if (vm->cursor < it.hi) nxt = std::min(nxt, (int)((ev[vm->cursor].time - t0) >> 8));
                if (nxt <= sfr) nxt = sfr + 1;      // defensive: never stall
                if (vm->alive)
                    for (auto &om : vm->osc) {
                        bool is_noise;
                        int draws = om.second.segment(e, nxt - sfr, &is_noise);
                        if (is_noise) {
                            HostEvent sd;
                            sd.time = t0 + ((uint64_t)sfr << 8);
                            sd.seq = 0xf0000000u + (uint32_t)om.first;  // after the writes of this frame
                            sd.voice = it.slot;
                            sd.y = EV_SEED | ((uint32_t)om.first << 8);
                            sd.value = (int)*e->noise_ptr;
                            sd.dur = 0;
                            seeds[it.bank].push_back(sd);
                            lcg_advance(e->noise_ptr, draws);
                        }
                    }
                sfr = nxt;
            }
            (void)b;
        }
        a = pe;
    }
    // 4. merge the seed records, drop mirrors that are no longer needed
    for (size_t bi = 0; bi < e->banks.size(); ++bi) {
        if (seeds[bi].empty()) continue;
        due[bi].insert(due[bi].end(), seeds[bi].begin(), seeds[bi].end());
        std::sort(due[bi].begin(), due[bi].end(), [](const HostEvent &a, const HostEvent &c) {
            if (a.voice != c.voice) return a.voice < c.voice;
            if ((a.time >> 8) != (c.time >> 8)) return a.time < c.time;
            return a.seq < c.seq;
        });
    }
    for (auto it = e->mirrors.begin(); it != e->mirrors.end();) {
        bool keep = false;
        if (it->second.alive)
            for (auto &om : it->second.osc)
                if (om.second.mode == OSC_NOISE) keep = true;
        if (e->banks[(int)(it->first >> 32)]->dynamic) keep = true;
        if (keep) ++it; else it = e->mirrors.erase(it);
    }
    return A2CU_OK;
}

static int run_window(a2cu_engine *e, unsigned frames, unsigned buffer, int32_t *dev_out, bool allow_lag = false) {
    if (!frames) return A2CU_OK;
    if (!buffer) buffer = frames;
    const uint64_t t0 = e->now, t1 = e->now + ((uint64_t)frames << 8);
    // the engine-owned master block must hold the WHOLE window before it may be split below
    if (!dev_out) {
        int gr = grow_device(e, &e->d_master, &e->master_cap, (size_t)frames * 2, true);
        if (gr) return gr;
    }
    int splits[kMaxSplits], nsplits = 0;
    // one launch renders at most kSplitMaxWin frames (render_split keeps the CTA's bus sums of the
    // whole window in shared memory); longer windows are rendered as sub-windows, like windows with
    // too many root-level cuts for one launch
    if (frames > (unsigned)kSplitMaxWin || collect_splits(e, t0, t1, splits, &nsplits)) {
        // Too many root-level cuts for one launch: render the window as two sub-windows, each
        // writing its own part of the output block. Cut at a driver-buffer boundary, or - inside
        // one buffer - at a fragment boundary (fragments restart every 64 frames from the buffer
        // start, core.c:1964-1973, so both parts keep the original fragment grid).
        const int och = e->post_root ? e->channels : 2;
        int32_t *base = dev_out ? dev_out : e->d_master;
        unsigned h, b0, b1;
        if (frames > buffer) {
            h = ((frames + buffer - 1) / buffer / 2) * buffer;
            b0 = b1 = buffer;
        } else if (frames > (unsigned)kMaxFrag) {
            h = ((frames + kMaxFrag - 1) / kMaxFrag / 2) * kMaxFrag;
            b0 = h; b1 = frames - h;
        } else
            return fail(A2CU_EINVAL, "too many root events in one fragment%s");
        int r = run_window(e, h, b0, base, allow_lag);
        if (r) return r;
        return run_window(e, frames - h, b1, (int32_t *)((char *)base + (size_t)h * och * out_sample_bytes(e)), allow_lag);
    }
    int r = upload_waves(e);
    if (r) return r;

    const int W = (int)frames;
    const int nbus = 1 + e->ngroups;
    size_t acc_n = (size_t)nbus * W * 2;
    if (acc_n > e->acc_cap) {
        r = grow_device(e, &e->d_acc, &e->acc_cap, acc_n, true);
        if (r) return r;
        e->acc_clean_n = 0;
    }

    // ---- stage events (host -> pinned -> device) ----
    const double ws_t0 = now_us();
    size_t stage_bytes = 0;
    // per-bank events of this window. The vectors are deliberately short-lived: with many banks served
    // round-robin the allocator hands the same (cache-hot) block to every window, whereas per-bank
    // persistent buffers would cycle through tens of MB of cold host memory
    std::vector<std::vector<HostEvent>> due(e->banks.size());
    std::vector<uint8_t> fast(e->banks.size(), 0);     // bank's window consists of bulk writes only
    const bool planner = e->noise_seen || !e->mirrors.empty();
    for (size_t bi = 0; bi < e->banks.size(); ++bi) {
        Bank *b = e->banks[bi];
        if (!b->enabled) continue;
        if (!b->bulk.empty()) {
            bool all_bulk_due = true;
            for (auto &be : b->bulk) if (be.time >= t1) all_bulk_due = false;
            if (b->events.empty() && all_bulk_due && !planner && !b->dynamic) {
                fast[bi] = 1;
                std::stable_sort(b->bulk.begin(), b->bulk.end(), [](const Bank::BulkEvent &a, const Bank::BulkEvent &c) {
                    if ((a.time >> 8) != (c.time >> 8)) return a.time < c.time;
                    return a.seq < c.seq;
                });
                continue;
            }
            // general path: bulk writes become ordinary per-voice events
            for (auto &be : b->bulk)
                for (int i = 0; i < b->nvoices; ++i) {
                    HostEvent ev;
                    ev.time = be.time; ev.seq = be.seq + (uint32_t)i; ev.voice = i; ev.y = be.y;
                    ev.value = be.has_values ? be.values[i] : be.scalar; ev.dur = be.dur;
                    b->events.push_back(ev);
                }
            b->bulk.clear();
        }
        if (b->events.empty()) continue;
        bool all_due = true;
        for (auto &ev : b->events)
            if (ev.time >= t1) { all_due = false; break; }
        if (all_due) {
            due[bi].swap(b->events);        // b->events is left empty and without storage
        } else {
            std::vector<HostEvent> keep;
            for (auto &ev : b->events)
                (ev.time < t1 ? due[bi] : keep).push_back(ev);
            b->events.swap(keep);
        }
        {
            auto less = [](const HostEvent &a, const HostEvent &c) {
                if (a.voice != c.voice) return a.voice < c.voice;
                if ((a.time >> 8) != (c.time >> 8)) return a.time < c.time;
                return a.seq < c.seq;
            };
            // bulk writes (a2cu_bank_write_all) arrive in voice order already
            if (!std::is_sorted(due[bi].begin(), due[bi].end(), less))
                std::sort(due[bi].begin(), due[bi].end(), less);
        }
        if (!due[bi].empty())
            stage_bytes += (b->stride + 1) * sizeof(unsigned) + due[bi].size() * sizeof(uint4) + 64;
    }
    if (planner) {
        r = plan_noise(e, t0, W, (int)buffer, splits, nsplits, due);
        if (r) return r;
    }
    stage_bytes = 0;
    for (size_t bi = 0; bi < e->banks.size(); ++bi) {
        const size_t nev = fast[bi] ? e->banks[bi]->bulk.size() * (size_t)e->banks[bi]->nvoices : due[bi].size();
        if (nev) stage_bytes += (e->banks[bi]->stride + 1) * sizeof(unsigned) + nev * sizeof(uint4) + 64;
    }
    std::vector<MixHostEvent> mdue, mkeep;
    for (auto &m : e->mixev) (m.time < t1 ? mdue : mkeep).push_back(m);
    e->mixev.swap(mkeep);
    std::sort(mdue.begin(), mdue.end(), [](const MixHostEvent &a, const MixHostEvent &c) {
        if ((a.time >> 8) != (c.time >> 8)) return a.time < c.time;
        return a.seq < c.seq;
    });
    stage_bytes += mdue.size() * sizeof(MixEvent) + 64;
    const double ws_t1 = now_us();
    r = ensure_stage(e, stage_bytes);
    if (r) return r;
    char *stage = (char *)e->h_stage;
    // bank event uploads use the copy stream unless bus-stage events share this staging buffer
    e->copy_used = mdue.empty() && !e->env_no_copy_stream;
    if (e->copy_used && !e->copy_stream) CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    cudaStream_t up = e->copy_used ? e->copy_stream : e->stream;
    size_t spos = 0;

    // the bus stage of the previous window zeroed the rows it read (MixParams::clear); a memset is only
    // needed when the layout [bus][W][2] changed or nothing has run yet
    if (e->acc_clean_n != acc_n || e->acc_clean_W != W) {
        CK(cudaMemsetAsync(e->d_acc, 0, acc_n * sizeof(int), e->stream));
    }
    e->acc_clean_n = acc_n; e->acc_clean_W = W;

    std::vector<RenderParams> &params = e->params;
    if (params.size() < e->banks.size()) params.resize(e->banks.size());
    for (size_t bi = 0; bi < e->banks.size(); ++bi) {
        Bank *b = e->banks[bi];
        if (b->dynamic || !b->nvoices || !b->enabled) continue;
        RenderParams &P = params[bi];
        memset(&P, 0, sizeof(P));
        P.state = b->d_state; P.stride = b->stride; P.nvoices = b->nvoices;
        P.bus_of = b->d_bus; P.acc = e->d_acc; P.W = W; P.buffer = (int)buffer;
        P.nsplits = nsplits;
        for (int i = 0; i < nsplits; ++i) P.splits[i] = splits[i];
        P.waves = e->d_waves; P.pool = e->d_pool; P.cpool = e->d_cpool; P.ptab = e->d_ptab; P.fmsine = e->d_fmsine;
        P.f12tab = e->d_f12tab;
        P.samplerate = e->samplerate;
        const size_t nbulk = fast[bi] ? b->bulk.size() : 0;
        if (!due[bi].empty() || nbulk) {
            size_t nev = nbulk ? nbulk * (size_t)b->nvoices : due[bi].size();
            // device layout: [CSR offsets (stride + 1, padded to 16 B) | records] - one H2D copy
            const size_t off_bytes = (((b->stride + 1) * sizeof(unsigned)) + 15) & ~(size_t)15;
            if (nev > b->ev_cap) {
                if (b->d_ev) cudaFree(b->d_ev);
                b->ev_cap = nev * 2;
                CK(cudaMalloc(&b->d_ev, off_bytes + b->ev_cap * sizeof(uint4)));
            }
            spos = (spos + 15) & ~(size_t)15;
            unsigned *off = (unsigned *)(stage + spos);
            uint4 *recs = (uint4 *)(stage + spos + off_bytes);
            spos += off_bytes + nev * sizeof(uint4);
            if (nbulk) {
                // every voice has the same nbulk records (already in time / sequence order)
                const unsigned K = (unsigned)nbulk;
                for (size_t v = 0; v <= b->stride; ++v) off[v] = (unsigned)std::min(v, (size_t)b->nvoices) * K;
                for (unsigned k = 0; k < K; ++k) {
                    const Bank::BulkEvent &be = b->bulk[k];
                    const unsigned rel = be.time >= t0 ? (unsigned)(be.time - t0) : (unsigned)(be.time & 0xff);
                    uint4 rec = make_uint4(rel, be.y, (unsigned)be.scalar, be.dur);
                    uint4 *dst = recs + k;
                    if (be.has_values) {
                        const int32_t *val = be.values.data();
                        for (int v = 0; v < b->nvoices; ++v, dst += K) { rec.z = (unsigned)val[v]; *dst = rec; }
                    } else
                        for (int v = 0; v < b->nvoices; ++v, dst += K) *dst = rec;
                }
            } else {
                // due[bi] is sorted by voice: CSR offsets in one pass
                size_t k = 0;
                const HostEvent *d = due[bi].data();
                for (size_t v = 0; v <= b->stride; ++v) {
                    while (k < nev && (size_t)d[k].voice < v) ++k;
                    off[v] = (unsigned)k;
                }
                for (size_t i = 0; i < nev; ++i) {
                    // events of a bank that was paused carry times before t0: they apply at the window start
                    const unsigned rel = d[i].time >= t0 ? (unsigned)(d[i].time - t0) : (unsigned)(d[i].time & 0xff);
                    recs[i] = make_uint4(rel, d[i].y, (unsigned)d[i].value, d[i].dur);
                }
            }
            if (e->copy_used && b->ev_consumed) CK(cudaStreamWaitEvent(e->copy_stream, b->ev_consumed, 0));
            CK(cudaMemcpyAsync(b->d_ev, off, off_bytes + nev * sizeof(uint4), cudaMemcpyHostToDevice, up));
            e->h2d_bytes += (b->stride + 1) * sizeof(unsigned) + nev * sizeof(uint4);
            b->d_evoff = (unsigned *)b->d_ev;
            b->d_evrecs = (uint4 *)((char *)b->d_ev + off_bytes);
            P.ev_off = b->d_evoff; P.ev = b->d_evrecs;
        }
    }
    if (e->copy_used) {
        // uploads done -> staging buffer reusable; the kernels wait for exactly this event
        const int k = (e->stage_pos + a2cu_engine::kStageRing - 1) % a2cu_engine::kStageRing;
        CK(cudaEventRecord(e->stage_done[k], e->copy_stream));
        CK(cudaStreamWaitEvent(e->stream, e->stage_done[k], 0));
    }
    // Fused root stage: when exactly one bank renders with render_split, there are no group buses, no
    // root events in this window, the last CTA of
    // that kernel runs the root stage itself (a2cu_split.cuh) and mix_root is not launched.
    int fuse_bank = -1;
    if (e->ngroups == 0 && mdue.empty() && e->use_split && !e->env_no_fuse) {
        int live = 0;
        for (size_t bi = 0; bi < e->banks.size(); ++bi) {
            Bank *b = e->banks[bi];
            if (b->dynamic || !b->nvoices || !b->enabled) continue;
            ++live;
            fuse_bank = (b->k.split[0].fn && !b->exotic) ? (int)bi : -1;
        }
        if (live != 1) fuse_bank = -1;
        if (fuse_bank >= 0 && !e->d_fuse_counter) {
            CK(cudaMalloc(&e->d_fuse_counter, sizeof(unsigned)));
            CK(cudaMemsetAsync(e->d_fuse_counter, 0, sizeof(unsigned), e->stream));
        }
    }
    e->fused_root = false;
    bool lag_consumed = false, lag_created = false;
    if (e->xchg.enabled && e->xchg.world > 1 && W > e->xchg.max_frames)
        return fail(A2CU_EINVAL, "window longer than the exchange buffer (a2cu_xchg_create max_frames)%s");
    const XchgParams X = xchg_params(e);
    const double ws_t2 = now_us();
    // inputs are resident from here on: ev0 .. ev1 brackets the render kernels
    if (e->timing) CK(cudaEventRecord(e->ev0, e->stream));
    int nlive = 0, launched = 0;
    for (Bank *b : e->banks)
        if (!b->dynamic && b->nvoices && b->enabled) ++nlive;
    const bool fan_out = nlive > 1 && !e->env_no_side_streams;
    bool side_used[a2cu_engine::kSideStreams] = {false, false, false};
    if (fan_out) {
        if (!e->side_fork) {
            CK(cudaEventCreateWithFlags(&e->side_fork, cudaEventDisableTiming));
            for (int k = 0; k < a2cu_engine::kSideStreams; ++k) {
                CK(cudaStreamCreateWithFlags(&e->side[k], cudaStreamNonBlocking));
                CK(cudaEventCreateWithFlags(&e->side_join[k], cudaEventDisableTiming));
            }
        }
        CK(cudaEventRecord(e->side_fork, e->stream));
    }
    for (size_t bi = 0; bi < e->banks.size(); ++bi) {
        Bank *b = e->banks[bi];
        if (b->dynamic || !b->nvoices || !b->enabled) continue;
        // bank k of the window runs on stream k mod 4: the engine's own stream or a side stream
        cudaStream_t ls = e->stream;
        const int lane = launched++ % (a2cu_engine::kSideStreams + 1);
        if (fan_out && lane > 0) {
            ls = e->side[lane - 1];
            if (!side_used[lane - 1]) {
                CK(cudaStreamWaitEvent(ls, e->side_fork, 0));
                side_used[lane - 1] = true;
            }
        }
        bool split = e->use_split && b->k.split[0].fn && !b->exotic && nsplits <= 1;
        // Large filtered banks: once every SM holds several hundred voices, one thread per voice
        // (render_bank) keeps more recurrences in flight per SM than the warp-specialised kernel
        // (profiles/r02_saturation.json: the curves cross between 32 k and 64 k voices on 148 SMs)
        if (split && b->nvoices > e->sm_count * 288 && !e->env_split_always) {
            for (const a2cu_unitspec &u : b->chain)
                if (u.kind == A2CU_FILTER12) split = false;
        }
        std::vector<HostEvent> bulk_probe;      // fast path: all voices share voice 0's event times
        if (fast[bi])
            for (auto &be : b->bulk) {
                HostEvent ev;
                ev.time = be.time; ev.seq = be.seq; ev.voice = 0; ev.y = be.y; ev.value = be.scalar; ev.dur = be.dur;
                bulk_probe.push_back(ev);
            }
        const std::vector<HostEvent> &scan = fast[bi] ? bulk_probe : due[bi];
        if (split && (!scan.empty() || nsplits)) {
            // at most kSplitSegs segments per voice and fragment
            auto frag_start = [&](int f) { int pos = f % (int)buffer; return f - pos + (pos / kMaxFrag) * kMaxFrag; };
            int split_frag = nsplits ? frag_start(splits[0]) : -1;
            if (nsplits && splits[0] == split_frag) split_frag = -1;    // on a fragment boundary: no extra cut
            int cur_voice = -1, cur_frag = -1, cnt = 0, last = -1;
            uint64_t memo_time = ~(uint64_t)0;
            int f = 0, fs = 0;
            for (const HostEvent &ev : scan) {
                if (ev.time != memo_time) {     // bulk writes share one time stamp
                    memo_time = ev.time;
                    f = ev.time >= t0 ? (int)((ev.time - t0) >> 8) : 0;
                    fs = frag_start(f);
                }
                if (ev.voice != cur_voice || fs != cur_frag) {
                    cur_voice = ev.voice; cur_frag = fs; last = -1;
                    cnt = (fs == split_frag) ? 1 : 0;
                }
                if (f != fs && f != last && !(fs == split_frag && f == splits[0])) { ++cnt; last = f; }
                if (cnt > kSplitSegs - 1) { split = false; break; }
            }
        }
        if (split) {
            params[bi].prof = e->d_prof;
            // two voice sets per CTA once the bank has more 32-voice sets than the chip has SMs
            int var = (b->k.split[1].fn && b->nvoices > e->sm_count * 32 && !e->env_one_set) ? 1 : 0;
            if (b->raw_taps) var = 2;
            size_t smem = b->k.split[var].smem;
            if (b->stage_wave >= 0 && e->waves[b->stage_wave].cbegin >= 0 && !e->env_no_stage) {
                // whole wave (all mip levels) + read-ahead slack, if it fits beside the pipeline buffers
                const HostWave &hw = e->waves[b->stage_wave];
                size_t tb = ((size_t)hw.ccount + 64) * sizeof(int4);
                if (smem + tb > kMaxSplitSmem && var == 1) { var = 0; smem = b->k.split[0].smem; }
                if (smem + tb <= kMaxSplitSmem) {
                    params[bi].stage_begin = hw.cbegin;
                    params[bi].stage_count = hw.ccount + 64;
                    smem += tb;
                }
            }
            const SplitVariant &sv = b->k.split[var];
            int grid = (b->nvoices + sv.voices - 1) / sv.voices;
            if ((int)bi == fuse_bank) {
                params[bi].fuse_root = 1;
                params[bi].fuse_counter = e->d_fuse_counter;
                params[bi].fuse_rstate = e->d_rstate;
                params[bi].fuse_master = dev_out ? dev_out : e->d_master;
                params[bi].fuse_channels = e->channels;
                params[bi].fuse_root_stage = e->post_root ? 1 : 0;
                params[bi].fuse_out_fmt = e->post_root ? e->out_fmt : 0;
                params[bi].xchg = X;
                if (X.world > 1 && allow_lag) {
                    // pipelined + sharded: this launch publishes its bus and finishes the pending window
                    params[bi].xchg.lag = 1;
                    xchg_fill_prev(e, params[bi].xchg);
                    lag_consumed = e->xchg.pending.valid;
                    a2cu_engine::Xchg::Pending &pn = e->xchg.pending;
                    pn.valid = true; pn.epoch = X.epoch; pn.W = W; pn.buffer = (int)buffer; pn.nsplits = nsplits;
                    for (int i = 0; i < nsplits; ++i) pn.splits[i] = splits[i];
                    pn.master = params[bi].fuse_master;
                    lag_created = true;
                } else if (e->xchg.pending.valid) {
                    r = xchg_drain_pending(e);
                    if (r) return r;
                }
                e->fused_root = true;
            }
            sv.fn<<<grid, sv.threads, smem, ls>>>(params[bi]);
            ++e->split_launches;
            if ((int)bi == fuse_bank && lag_consumed) {
                r = xchg_pending_done(e);       // the previous window's output is complete after this launch
                if (r) return r;
            }
        } else {
            int grid = (b->nvoices + kThreads - 1) / kThreads;
            if (b->generic) render_generic<<<grid, kThreads, 0, ls>>>(params[bi], b->gchain, b->d_scratch);
            else b->k.fn<<<grid, kThreads, 0, ls>>>(params[bi]);
        }
        ++e->launches;
        if (!b->ev_consumed) CK(cudaEventCreateWithFlags(&b->ev_consumed, cudaEventDisableTiming));
        CK(cudaEventRecord(b->ev_consumed, ls));
        if (fast[bi]) b->bulk.clear();      // consumed (staged above)
    }
    for (int k = 0; k < a2cu_engine::kSideStreams; ++k)
        if (side_used[k]) {                 // join: the bus stage needs every bank's sums
            CK(cudaEventRecord(e->side_join[k], e->side[k]));
            CK(cudaStreamWaitEvent(e->stream, e->side_join[k], 0));
        }
    if (e->timing) CK(cudaEventRecord(e->ev1, e->stream));
    const double ws_t3 = now_us();

    MixParams M;
    memset(&M, 0, sizeof(M));
    M.acc = e->d_acc; M.W = W; M.buffer = (int)buffer; M.ngroups = e->ngroups; M.channels = e->channels;
    M.nsplits = nsplits;
    for (int i = 0; i < nsplits; ++i) M.splits[i] = splits[i];
    M.gstate = e->d_gstate; M.rstate = e->d_rstate;
    M.nev = (int)mdue.size();
    if (M.nev) {
        if (mdue.size() > e->mixev_cap) {
            if (e->d_mixev) cudaFree(e->d_mixev);
            e->mixev_cap = mdue.size() * 2;
            CK(cudaMalloc(&e->d_mixev, e->mixev_cap * sizeof(MixEvent)));
        }
        spos = (spos + 15) & ~(size_t)15;
        MixEvent *me = (MixEvent *)(stage + spos);
        for (size_t i = 0; i < mdue.size(); ++i) {
            uint64_t rel = mdue[i].time >= t0 ? mdue[i].time - t0 : (mdue[i].time & 0xff);
            me[i].time = (unsigned)rel; me[i].target = mdue[i].target;
            me[i].reg_dur_hi = mdue[i].reg & 0xff; me[i].value = mdue[i].value; me[i].dur = mdue[i].dur;
        }
        CK(cudaMemcpyAsync(e->d_mixev, me, mdue.size() * sizeof(MixEvent), cudaMemcpyHostToDevice, e->stream));
        e->h2d_bytes += mdue.size() * sizeof(MixEvent);
        M.ev = e->d_mixev;
    }
    if (!e->copy_used) {
        r = stage_release(e);
        if (r) return r;
    }
    M.master = dev_out ? dev_out : e->d_master;
    M.root_stage = e->post_root ? 1 : 0;
    M.out_fmt = e->post_root ? e->out_fmt : 0;
    M.clear = 1;
    if (e->root_written && t0 < e->root_until) { M.general = 1; e->root_extra = true; }
    else if (e->root_extra) { M.general = 1; e->root_extra = false; }
    if (e->ngroups) {
        mix_groups<<<e->ngroups, 256, 0, e->stream>>>(M);
        ++e->launches;
    }
    if (!e->fused_root && e->xchg.pending.valid) {
        r = xchg_drain_pending(e);      // root stages run in window order (the rampers carry over)
        if (r) return r;
    }
    (void)lag_created;
    if (!e->fused_root) {
        if (X.world > 1) mix_root_xchg<<<1, 512, 0, e->stream>>>(M, X);
        else mix_root<<<M.general ? 1 : std::max(1, std::min(8, ((int)M.W + 255) / 256)), 256, 0, e->stream>>>(M);
        ++e->launches;
    }
    CK(cudaGetLastError());
    if (e->timing) CK(cudaEventRecord(e->ev2, e->stream));
    e->last_mix = M;
    e->now = t1;
    e->ws.prep_us += ws_t1 - ws_t0; e->ws.stage_us += ws_t2 - ws_t1; e->ws.api_us += ws_t3 - ws_t2;
    e->ws.mix_us += now_us() - ws_t3; ++e->ws.windows;
    return A2CU_OK;
}

static int xchg_pending_done(a2cu_engine *e) {
    const int k = e->xchg.deferred_slot;
    if (k >= 0) {
        CK(cudaEventRecord(e->slots[k].done, e->stream));
        e->xchg.deferred_slot = -1;
    }
    return A2CU_OK;
}
static int xchg_drain_pending(a2cu_engine *e) {
    if (!e->xchg.pending.valid) return A2CU_OK;
    XchgParams X;
    memset(&X, 0, sizeof(X));
    a2cu_engine::Xchg &x = e->xchg;
    X.world = x.world; X.rank = x.rank; X.max_frames = x.max_frames;
    for (int r = 0; r < x.world; ++r) {
        X.flags[r] = (unsigned *)x.peer[r];
        X.data[r] = (int *)((char *)x.peer[r] + kXchgFlagBytes);
    }
    X.status = x.h_status; X.timeout_cycles = x.timeout_cycles;
    xchg_fill_prev(e, X);
    xchg_drain<<<1, 512, 0, e->stream>>>(X, e->d_rstate, e->channels, e->post_root ? 1 : 0, e->post_root ? e->out_fmt : 0);
    ++e->launches;
    CK(cudaGetLastError());
    x.pending.valid = false;
    return xchg_pending_done(e);
}

int a2cu_run_async(a2cu_engine *e, unsigned frames, unsigned buffer, int32_t *dev_out) {
    if (!e) return A2CU_EINVAL;
    cudaSetDevice(e->device);
    return run_window(e, frames, buffer, dev_out);
}

int32_t *a2cu_master_devptr(a2cu_engine *e) { return e ? e->d_master : nullptr; }

int a2cu_sync(a2cu_engine *e) {
    if (!e) return A2CU_EINVAL;
    if (e->xchg.pending.valid) {
        cudaSetDevice(e->device);
        int r = xchg_drain_pending(e);
        if (r) return r;
    }
    CK(cudaStreamSynchronize(e->stream));
    if (e->timing) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) e->last_ms = ms;
        if (cudaEventElapsedTime(&ms, e->ev1, e->ev2) == cudaSuccess) e->last_mix_ms = ms;
    }
    return xchg_check(e);
}

int a2cu_run(a2cu_engine *e, unsigned frames, unsigned buffer, int32_t *out) {
    if (!e) return A2CU_EINVAL;
    cudaSetDevice(e->device);
    const int och = e->post_root ? e->channels : 2;
    size_t n = (size_t)frames * och;
    int r = run_window(e, frames, buffer, nullptr);
    if (r) return r;
    if (out) {
        if (n > e->hout_cap) {
            if (e->h_out) cudaFreeHost(e->h_out);
            e->hout_cap = n * 2;
            CK(cudaMallocHost(&e->h_out, e->hout_cap * sizeof(int32_t)));
        }
        CK(cudaMemcpyAsync(e->h_out, e->d_master, n * out_sample_bytes(e), cudaMemcpyDeviceToHost, e->stream));
        e->d2h_bytes += n * out_sample_bytes(e);
    }
    r = a2cu_sync(e);
    if (r) return r;
    if (out) memcpy(out, e->h_out, n * out_sample_bytes(e));
    return A2CU_OK;
}

// ---- pipelined rendering ---------------------------------------------------
// a2cu_submit queues one window (event staging, H2D, kernels, D2H of the
// master block into a pinned result slot) and returns at once; a2cu_collect
// waits for that window only. With two windows in flight the host stages
// window i+1 while the device renders window i.
static int submit_impl(a2cu_engine *e, unsigned frames, unsigned buffer, int32_t *dev_out);
int a2cu_submit(a2cu_engine *e, unsigned frames, unsigned buffer) { return submit_impl(e, frames, buffer, nullptr); }
// Same, but the window's output (master block, or the raw root bus with post_root_stage 0) goes to
// DEVICE memory 'dev_out' - for callers that go on with it on the stream (the multi-GPU reduce).
// a2cu_collect(ticket, NULL) then only waits for the window's kernels and latches its timing spans.
int a2cu_submit_dev(a2cu_engine *e, unsigned frames, unsigned buffer, int32_t *dev_out) {
    if (!dev_out) return fail(A2CU_EINVAL, "a2cu_submit_dev: dev_out is NULL%s");
    return submit_impl(e, frames, buffer, dev_out);
}
static int submit_impl(a2cu_engine *e, unsigned frames, unsigned buffer, int32_t *dev_out) {
    if (!e) return A2CU_EINVAL;
    cudaSetDevice(e->device);
    const int k = e->slot_pos;
    a2cu_engine::Slot &sl = e->slots[k];
    if (sl.busy) return fail(A2CU_EINVAL, "a2cu_submit: %s", "all result slots in flight, a2cu_collect first");
    if (!sl.done) {
        CK(cudaEventCreate(&sl.ev0)); CK(cudaEventCreate(&sl.ev1)); CK(cudaEventCreate(&sl.ev2));
        CK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    }
    const int och = e->post_root ? e->channels : 2;
    const size_t n = (size_t)frames * och;
    if (!dev_out && n > sl.cap) {
        if (sl.h_out) cudaFreeHost(sl.h_out);
        sl.cap = n * 2;
        CK(cudaMallocHost(&sl.h_out, sl.cap * sizeof(int32_t)));
    }
    // this window's spans go to the slot's own events
    cudaEvent_t s0 = e->ev0, s1 = e->ev1, s2 = e->ev2;
    e->ev0 = sl.ev0; e->ev1 = sl.ev1; e->ev2 = sl.ev2;
    // the bus stage writes the master block straight into the pinned result slot (mapped host memory,
    // same address on the device under UVA): no separate D2H copy on the stream
    int r = run_window(e, frames, buffer, dev_out ? dev_out : sl.h_out, /*allow_lag=*/true);
    e->ev0 = s0; e->ev1 = s1; e->ev2 = s2;
    if (r) return r;
    // sharded + lagged: this window's output is finished by the next launch (or by a2cu_collect)
    if (e->xchg.pending.valid) e->xchg.deferred_slot = k;
    else CK(cudaEventRecord(sl.done, e->stream));
    if (!dev_out) e->d2h_bytes += n * out_sample_bytes(e);
    sl.n = dev_out ? 0 : n;
    sl.busy = true;
    e->slot_pos = (k + 1) % a2cu_engine::kSlots;
    return k;
}

int a2cu_collect(a2cu_engine *e, int ticket, int32_t *out) {
    if (!e || ticket < 0 || ticket >= a2cu_engine::kSlots || !e->slots[ticket].busy)
        return fail(A2CU_EINVAL, "a2cu_collect: bad ticket%s");
    cudaSetDevice(e->device);
    a2cu_engine::Slot &sl = e->slots[ticket];
    if (e->xchg.deferred_slot == ticket) {      // nothing was submitted after it: finish it now
        int r = xchg_drain_pending(e);
        if (r) return r;
    }
    CK(cudaEventSynchronize(sl.done));
    if (out && sl.n) memcpy(out, sl.h_out, sl.n * out_sample_bytes(e));
    if (e->timing) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sl.ev0, sl.ev1) == cudaSuccess) e->last_ms = ms;
        if (cudaEventElapsedTime(&ms, sl.ev1, sl.ev2) == cudaSuccess) e->last_mix_ms = ms;
    }
    sl.busy = false;
    return xchg_check(e);
}

int a2cu_apply_root_stage(a2cu_engine *e, const int32_t *dev_rootbus, int32_t *dev_master, unsigned frames,
                          unsigned buffer, uint64_t start_time) {
    // Runs only the root panmix over an externally summed root bus.
    if (!e || !dev_rootbus || !dev_master) return A2CU_EINVAL;
    cudaSetDevice(e->device);
    MixParams M = e->last_mix;
    (void)start_time;
    M.acc = (int *)dev_rootbus; M.W = (int)frames; M.buffer = (int)(buffer ? buffer : frames);
    M.ngroups = 0; M.master = dev_master; M.root_stage = 1; M.clear = 0;
    mix_root<<<M.general ? 1 : std::max(1, std::min(8, ((int)M.W + 255) / 256)), 256, 0, e->stream>>>(M);
    ++e->launches;
    CK(cudaGetLastError());
    return A2CU_OK;
}

// Device-side f12_pitch2coeff for an array of cutoff ramper values (8:24), for the parity tests.
int a2cu_debug_f12_coeff(a2cu_engine *e, const int32_t *cutoff_values, int n, int32_t *out) {
    if (!e || !cutoff_values || !out || n < 1) return A2CU_EINVAL;
    cudaSetDevice(e->device);
    int *d = nullptr;
    CK(cudaMalloc(&d, (size_t)n * 2 * sizeof(int)));
    cudaError_t err = memcpy_done(d, cutoff_values, (size_t)n * sizeof(int), cudaMemcpyHostToDevice);
    Ctx c;
    memset(&c, 0, sizeof(c));
    c.ptab = e->d_ptab; c.f12tab = e->d_f12tab; c.samplerate = e->samplerate;
    if (err == cudaSuccess) {
        f12_coeff_probe<<<(n + 255) / 256, 256, 0, e->stream>>>(c, d, n, d + n);
        err = cudaStreamSynchronize(e->stream);
    }
    if (err == cudaSuccess) err = memcpy_done(out, d + n, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (err != cudaSuccess) return fail(A2CU_ECUDA, "a2cu_debug_f12_coeff: %s", cudaGetErrorString(err));
    return A2CU_OK;
}

// ===========================================================================
// Multi-GPU: root-bus exchange over NVLink peer memory (include/a2cu.h)
// ===========================================================================
int a2cu_xchg_create(a2cu_engine *e, int rank, int world, unsigned max_frames, unsigned timeout_ms,
                     void *handle_out) {
    if (!e || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || !max_frames)
        return fail(A2CU_EINVAL, "a2cu_xchg_create: bad args%s");
    cudaSetDevice(e->device);
    CK(cudaStreamSynchronize(e->stream));
    xchg_release(e);
    a2cu_engine::Xchg &x = e->xchg;
    x.world = world; x.rank = rank; x.max_frames = (int)max_frames;
    const size_t bytes = kXchgFlagBytes + (size_t)2 * world * max_frames * 2 * sizeof(int);
    if (cudaMalloc(&x.base, bytes) != cudaSuccess)
        return fail(A2CU_ENOMEM, "cudaMalloc exchange buffer: %s", cudaGetErrorString(cudaGetLastError()));
    CK(memset_done(x.base, 0, bytes));
    CK(cudaMalloc(&x.d_sum, (size_t)max_frames * 2 * sizeof(int)));
    CK(cudaHostAlloc((void **)&x.h_status, sizeof(unsigned), cudaHostAllocMapped));
    *x.h_status = 0;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, e->device);
    if (khz <= 0) khz = 1900000;
    x.timeout_cycles = (unsigned long long)(timeout_ms ? timeout_ms : 2000) * (unsigned long long)khz;
    x.peer[rank] = x.base;
    if (handle_out) {
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, x.base));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(handle_out, &h, sizeof(h));
    }
    return A2CU_OK;
}

int a2cu_xchg_connect_ipc(a2cu_engine *e, const void *handles) {
    if (!e || !handles || !e->xchg.base) return fail(A2CU_EINVAL, "a2cu_xchg_connect_ipc: create first%s");
    cudaSetDevice(e->device);
    a2cu_engine::Xchg &x = e->xchg;
    for (int r = 0; r < x.world; ++r) {
        if (r == x.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess) return fail(A2CU_ECUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(err));
        x.peer[r] = p; x.ipc_opened[r] = true;
    }
    x.enabled = true;
    return A2CU_OK;
}

int a2cu_xchg_connect_local(a2cu_engine *e, a2cu_engine *const *peers) {
    if (!e || !peers || !e->xchg.base) return fail(A2CU_EINVAL, "a2cu_xchg_connect_local: create first%s");
    cudaSetDevice(e->device);
    a2cu_engine::Xchg &x = e->xchg;
    for (int r = 0; r < x.world; ++r) {
        if (r == x.rank) continue;
        const a2cu_engine *p = peers[r];
        if (!p || !p->xchg.base || p->xchg.world != x.world || p->xchg.rank != r ||
            p->xchg.max_frames != x.max_frames)
            return fail(A2CU_EINVAL, "a2cu_xchg_connect_local: peer %s", "not created with the same geometry");
        if (p->device != e->device) {
            cudaError_t err = cudaDeviceEnablePeerAccess(p->device, 0);
            if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled)
                return fail(A2CU_ECUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(err));
            cudaGetLastError();
        }
        x.peer[r] = p->xchg.base;
    }
    x.enabled = true;
    return A2CU_OK;
}

int a2cu_xchg_enable(a2cu_engine *e, int enabled) {
    if (!e) return A2CU_EINVAL;
    if (!enabled && e->xchg.pending.valid) {
        cudaSetDevice(e->device);
        int r = xchg_drain_pending(e);
        if (r) return r;
    }
    if (enabled && !e->xchg.base) return fail(A2CU_EINVAL, "a2cu_xchg_enable: no exchange buffer%s");
    if (enabled)
        for (int r = 0; r < e->xchg.world; ++r)
            if (!e->xchg.peer[r]) return fail(A2CU_EINVAL, "a2cu_xchg_enable: peers not connected%s");
    e->xchg.enabled = enabled != 0;
    return A2CU_OK;
}

int a2cu_xchg_close(a2cu_engine *e) {
    if (!e) return A2CU_EINVAL;
    cudaSetDevice(e->device);
    if (e->xchg.pending.valid) xchg_drain_pending(e);
    cudaStreamSynchronize(e->stream);
    xchg_release(e);
    return A2CU_OK;
}

// ===========================================================================
// Drop-in ("block") mode: the API the unit plug-in (plugin/a2cu_units.c)
// records into while the reference host walks its voice tree.
// ===========================================================================
int a2cu_pool_open(a2cu_engine *e, const a2cu_unitspec *chain, int nunits) {
    if (!e || !chain || nunits < 1) return fail(A2CU_EINVAL, "a2cu_pool_open: bad args%s");
    std::string sig = sig_of(chain, nunits);
    for (size_t i = 0; i < e->banks.size(); ++i)
        if (e->banks[i]->dynamic && sig_of(e->banks[i]->chain.data(), (int)e->banks[i]->chain.size()) == sig)
            return (int)i;
    KernelEntry ke;
    GenericChain gc;
    bool generic = false;
    if (!find_kernel(chain, nunits, &ke, &gc, &generic))
        return fail(A2CU_ENOTIMPL, "no kernel for voice structure %s", sig.c_str());
    cudaSetDevice(e->device);
    Bank *b = new Bank();
    b->chain.assign(chain, chain + nunits);
    b->k = ke;
    b->generic = generic;
    b->gchain = gc;
    b->dynamic = true;
    b->nvoices = 0;
    b->stride = 1024;
    size_t sbytes = (size_t)b->k.words * b->stride * sizeof(int);
    if (cudaMalloc(&b->d_state, sbytes) != cudaSuccess || cudaMalloc(&b->d_noise, b->stride * sizeof(unsigned)) != cudaSuccess) {
        delete b;
        return fail(A2CU_ENOMEM, "cudaMalloc pool: %s", cudaGetErrorString(cudaGetLastError()));
    }
    CK(memset_done(b->d_state, 0, sbytes));
    CK(memset_done(b->d_noise, 0, b->stride * sizeof(unsigned)));
    if (b->generic) {
        CK(cudaMalloc(&b->d_scratch, b->stride * kMaxFrag * 2 * sizeof(int)));
        CK(memset_done(b->d_scratch, 0, b->stride * kMaxFrag * 2 * sizeof(int)));
    }
    e->banks.push_back(b);
    return (int)e->banks.size() - 1;
}

static int pool_grow(a2cu_engine *e, Bank *b) {
    size_t ns = b->stride * 2;
    int *nstate = nullptr;
    unsigned *nnoise = nullptr;
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMalloc(&nstate, (size_t)b->k.words * ns * sizeof(int)));
    CK(cudaMalloc(&nnoise, ns * sizeof(unsigned)));
    CK(memset_done(nstate, 0, (size_t)b->k.words * ns * sizeof(int)));
    CK(memset_done(nnoise, 0, ns * sizeof(unsigned)));
    CK(cudaMemcpy2D(nstate, ns * sizeof(int), b->d_state, b->stride * sizeof(int), b->stride * sizeof(int),
                    b->k.words, cudaMemcpyDeviceToDevice));
    cudaFree(b->d_state); cudaFree(b->d_noise);
    b->d_state = nstate; b->d_noise = nnoise; b->stride = ns;
    if (b->generic) {       // scratch rows hold nothing across fragments
        cudaFree(b->d_scratch);
        CK(cudaMalloc(&b->d_scratch, ns * kMaxFrag * 2 * sizeof(int)));
        CK(memset_done(b->d_scratch, 0, ns * kMaxFrag * 2 * sizeof(int)));
    }
    return A2CU_OK;
}

int a2cu_pool_alloc(a2cu_engine *e, int pool) {
    Bank *b = get_bank(e, pool);
    if (!b || !b->dynamic) return fail(A2CU_EINVAL, "bad pool%s");
    cudaSetDevice(e->device);
    if (!b->free_slots.empty()) {
        int s = b->free_slots.back();
        b->free_slots.pop_back();
        return s;
    }
    if ((size_t)b->used >= b->stride) {
        int r = pool_grow(e, b);
        if (r) return r;
    }
    return b->used++;
}

int a2cu_pool_free(a2cu_engine *e, int pool, int slot) {
    Bank *b = get_bank(e, pool);
    if (!b || !b->dynamic || slot < 0 || slot >= b->used) return fail(A2CU_EINVAL, "bad pool/slot%s");
    b->deferred_free.push_back(slot);     // still referenced by records of this block
    if (b->mirrored(slot)) { e->mirrors.erase(((uint64_t)pool << 32) | (uint32_t)slot); b->set_mirrored(slot, false); }
    return A2CU_OK;
}

int a2cu_block_begin(a2cu_engine *e) {
    if (!e) return A2CU_EINVAL;
    cudaSetDevice(e->device);
    int fr = a2cu_block_flush(e);           // nothing of the previous block may be left
    if (fr) return fr;
    for (Bank *b : e->banks) {
        b->free_slots.insert(b->free_slots.end(), b->deferred_free.begin(), b->deferred_free.end());
        b->deferred_free.clear();
    }
    e->pm_free.insert(e->pm_free.end(), e->pm_deferred.begin(), e->pm_deferred.end());
    e->pm_deferred.clear();
    for (int id : e->gunit_deferred)
        if (e->gunits[id].fbd) { e->fbd_free.push_back(e->gunits[id].fbd); e->gunits[id].fbd = nullptr; }
    e->gunit_free.insert(e->gunit_free.end(), e->gunit_deferred.begin(), e->gunit_deferred.end());
    e->gunit_deferred.clear();
    if (e->prev_nbbus < e->nbbus) e->prev_nbbus = e->nbbus;
    if (e->prev_nbbus && e->d_bacc)
        CK(cudaMemsetAsync(e->d_bacc, 0, (size_t)e->prev_nbbus * kMaxFrag * 2 * sizeof(int), e->stream));
    e->prev_nbbus = 0;
    e->nbbus = 0;
    e->buscmds.clear(); e->runs.clear(); e->cur_run = -1; ++e->flush_serial;
    for (Bank *b : e->banks) { b->bev.clear(); b->bruns.clear(); b->cur_slot = -1; }
    int r = upload_waves(e);
    return r;
}

int a2cu_block_bus(a2cu_engine *e) {
    if (!e) return A2CU_EINVAL;
    if (e->nbbus + 1 > e->bacc_cap) {
        int ncap = std::max(256, e->bacc_cap * 2);
        int *n = nullptr;
        CK(cudaStreamSynchronize(e->stream));
        CK(cudaMalloc(&n, (size_t)ncap * kMaxFrag * 2 * sizeof(int)));
        CK(memset_done(n, 0, (size_t)ncap * kMaxFrag * 2 * sizeof(int)));
        if (e->d_bacc) {
            CK(memcpy_done(n, e->d_bacc, (size_t)e->bacc_cap * kMaxFrag * 2 * sizeof(int), cudaMemcpyDeviceToDevice));
            cudaFree(e->d_bacc);
        }
        e->d_bacc = n;
        e->bacc_cap = ncap;
    }
    return e->nbbus++;
}

static inline void block_rec(Bank *b, int slot, unsigned x, unsigned y, int z, unsigned w) {
    if (b->cur_slot != slot || b->bruns.empty()) {
        VoiceRun r;
        r.slot = slot; r.ev_begin = (unsigned)b->bev.size(); r.ev_count = 0;
        b->bruns.push_back(r);
        b->cur_slot = slot;
        // a slot with two runs in one flush (its parent was cut by a wake-up): merged before the launch
        if (b->slot_mark.size() < b->stride) b->slot_mark.resize(b->stride, 0);
        if (b->slot_mark[slot] == b->flush_id) b->dup_runs = true;
        b->slot_mark[slot] = b->flush_id;
    }
    b->bev.push_back(make_uint4(x, y, (unsigned)z, w));
    ++b->bruns.back().ev_count;
}

int a2cu_block_init(a2cu_engine *e, int pool, int slot, int unit, int transpose, unsigned frame, unsigned substart) {
    Bank *b = get_bank(e, pool);
    if (!b || !b->dynamic || unit < 0 || unit >= (int)b->chain.size()) return fail(A2CU_EINVAL, "bad pool/unit%s");
    int kind = b->chain[unit].kind;
    int arg = init_arg(e, kind, transpose);
    unsigned x = (frame << 8) | (substart & 0xff);
    if (unit == 0 && b->mirrored(slot)) {       // slot reused by a new voice
        e->mirrors.erase(((uint64_t)pool << 32) | (uint32_t)slot);
        b->set_mirrored(slot, false);
    }
    block_rec(b, slot, x, EV_INIT | ((unsigned)unit << 8), arg, 0);
    if (kind == A2CU_FILTER12)
        block_rec(b, slot, x, EV_WRITE | ((unsigned)unit << 8) | (5u << 16),
                  tables().f12_coeff((int)((unsigned)arg << 8), e->samplerate), 0);
    if (kind == A2CU_DCBLOCK)
        block_rec(b, slot, x, EV_WRITE | ((unsigned)unit << 8),
                  tables().f12_coeff((int)((unsigned)(transpose - (5 << 16)) << 8), e->samplerate), 0);
    return A2CU_OK;
}

int a2cu_block_write(a2cu_engine *e, int pool, int slot, int unit, int reg, int32_t value, int transpose,
                     unsigned frame, unsigned start, uint32_t dur) {
    Bank *b = get_bank(e, pool);
    if (!b || !b->dynamic || unit < 0 || unit >= (int)b->chain.size()) return fail(A2CU_EINVAL, "bad pool/unit%s");
    Cooked c[2];
    int n = cook(e, b->chain[unit].kind, reg, value, (int)(start & 0xff), dur, transpose, c);
    if (n < 0) return n;
    unsigned x = (frame << 8) | (start & 0xff);
    if (b->chain[unit].kind == A2CU_WTOSC) {
        uint64_t key = ((uint64_t)pool << 32) | (uint32_t)slot;
        auto mi = b->mirrored(slot) ? e->mirrors.find(key) : e->mirrors.end();
        if (mi == e->mirrors.end() && c[0].reg == 0 && c[0].value >= 0 &&
            e->waves[c[0].value].type == A2CU_WNOISE) {
            // first noise selection of this voice: run what is recorded, then
            // read the oscillators' control state back once
            int r = a2cu_block_flush(e);
            if (r) return r;
            VoiceMirror *vm;
            r = mirror_create(e, pool, slot, &vm);
            if (r) return r;
            vm->alive = 1;
            mi = e->mirrors.find(key);
        }
        if (mi != e->mirrors.end())
            for (auto &om : mi->second.osc)
                if (om.first == unit) om.second.write(e, c[0].reg, c[0].value, (int)(start & 0xff), (int)c[0].dur);
    }
    for (int i = 0; i < n; ++i)
        block_rec(b, slot, x, EV_WRITE | ((unsigned)unit << 8) | ((unsigned)(c[i].reg & 0xff) << 16), c[i].value,
                  c[i].dur);
    return A2CU_OK;
}

int a2cu_block_proc(a2cu_engine *e, int pool, int slot, unsigned frame, unsigned frames, int bus) {
    Bank *b = get_bank(e, pool);
    if (!b || !b->dynamic || frames < 1 || frame + frames > (unsigned)kMaxFrag || bus < 0 || bus >= e->nbbus)
        return fail(A2CU_EINVAL, "a2cu_block_proc: bad args%s");
    auto mi = b->mirrored(slot) ? e->mirrors.find(((uint64_t)pool << 32) | (uint32_t)slot) : e->mirrors.end();
    if (mi != e->mirrors.end())
        for (auto &om : mi->second.osc) {      // units run in chain order inside one segment
            bool is_noise;
            int draws = om.second.segment(e, (int)frames, &is_noise);
            if (is_noise) {
                block_rec(b, slot, frame << 8, EV_SEED | ((unsigned)om.first << 8), (int)*e->noise_ptr, 0);
                lcg_advance(e->noise_ptr, draws);
            }
        }
    block_rec(b, slot, frame << 8, EV_PROC | (frames << 8), bus, 0);
    return A2CU_OK;
}

// Every bus command belongs to the run (voice) selected by a2cu_block_run().
static void push_cmd(a2cu_engine *e, BusCmd &c) {
    if (e->cur_run < 0) {       // caller did not name a voice: one run at the shallowest level
        a2cu_engine::HostRun r = {0, 0};
        e->runs.push_back(r);
        e->cur_run = (int)e->runs.size() - 1;
    }
    c.run = e->cur_run;
    ++e->runs[e->cur_run].count;
    e->buscmds.push_back(c);
}

uint64_t a2cu_block_run(a2cu_engine *e, int level, uint64_t prev) {
    if (!e) return 0;
    if ((uint32_t)(prev >> 32) == e->flush_serial && (uint32_t)prev < e->runs.size()) {
        e->cur_run = (int)(uint32_t)prev;
        return prev;
    }
    a2cu_engine::HostRun r = {level < 0 ? 0 : level, 0};
    e->runs.push_back(r);
    e->cur_run = (int)e->runs.size() - 1;
    return ((uint64_t)e->flush_serial << 32) | (uint32_t)e->cur_run;
}

int a2cu_pm_alloc(a2cu_engine *e) {
    if (!e) return A2CU_EINVAL;
    cudaSetDevice(e->device);
    int id;
    if (!e->pm_free.empty()) { id = e->pm_free.back(); e->pm_free.pop_back(); }
    else {
        if (e->pm_used + 1 > e->pm_cap) {
            int ncap = std::max(256, e->pm_cap * 2);
            int *n = nullptr;
            CK(cudaStreamSynchronize(e->stream));
            CK(cudaMalloc(&n, (size_t)ncap * 8 * sizeof(int)));
            if (e->d_pmstate) {
                CK(memcpy_done(n, e->d_pmstate, (size_t)e->pm_cap * 8 * sizeof(int), cudaMemcpyDeviceToDevice));
                cudaFree(e->d_pmstate);
            }
            e->d_pmstate = n; e->pm_cap = ncap;
        }
        id = e->pm_used++;
    }
    int st[8] = {65536 << 8, 65536 << 8, 0, 0, 0, 0, 0, 0};     // panmix.c:252-262
    CK(cudaMemcpyAsync(e->d_pmstate + (size_t)id * 8, st, sizeof(st), cudaMemcpyHostToDevice, e->stream));
    return id;
}

int a2cu_pm_free(a2cu_engine *e, int pm) {
    if (!e || pm < 0 || pm >= e->pm_used) return A2CU_EINVAL;
    e->pm_deferred.push_back(pm);
    return A2CU_OK;
}

int a2cu_block_pm_write(a2cu_engine *e, int pm, int reg, int32_t value, unsigned start, uint32_t dur) {
    if (!e || pm < 0 || pm >= e->pm_used || reg < 0 || reg > 1) return fail(A2CU_EINVAL, "bad panmix write%s");
    BusCmd c;
    memset(&c, 0, sizeof(c));
    c.op = BUS_PM_WRITE; c.pm = pm; c.reg = reg; c.value = value; c.start = (int)(start & 0xff); c.dur = (int)dur;
    push_cmd(e, c);
    return A2CU_OK;
}

int a2cu_block_pm_proc(a2cu_engine *e, int pm, int nin, int nout, int add, int in_bus, int out_bus, unsigned frame,
                       unsigned frames) {
    if (!e || pm < 0 || pm >= e->pm_used || in_bus < 0 || in_bus >= e->nbbus || out_bus < 0 || out_bus >= e->nbbus ||
        frames < 1 || frame + frames > (unsigned)kMaxFrag || nin < 1 || nin > 2 || nout < 1 || nout > 2)
        return fail(A2CU_EINVAL, "a2cu_block_pm_proc: bad args%s");
    BusCmd c;
    memset(&c, 0, sizeof(c));
    c.op = BUS_PM_PROC; c.pm = pm; c.nin = nin; c.nout = nout; c.add = add ? 1 : 0;
    c.in_bus = in_bus; c.out_bus = out_bus; c.frame = (int)frame; c.frames = (int)frames;
    push_cmd(e, c);
    return A2CU_OK;
}

// ---- generic units: one replaced unit, called on its own (BUS_U_* ops) -------
int a2cu_unit_alloc(a2cu_engine *e, int kind, int nin, int nout) {
    if (!e) return A2CU_EINVAL;
    a2cu_unitspec sp = {kind, nin, nout, 0, 0};
    bool known = kind == A2CU_WTOSC || kind == A2CU_PANMIX || kind == A2CU_FILTER12 || kind == A2CU_WAVESHAPER ||
                 kind == A2CU_FBDELAY || kind == A2CU_LIMITER || kind == A2CU_DCBLOCK || kind == A2CU_DC ||
                 (kind >= A2CU_FM1 && kind <= A2CU_FM4R);
    if (!known || nin < 0 || nin > 2 || nout < 1 || nout > 2 || unit_words(sp) > kUnitWords)
        return fail(A2CU_ENOTIMPL, "a2cu_unit_alloc: unsupported unit / channel count%s");
    if ((kind == A2CU_FILTER12 || kind == A2CU_WAVESHAPER || kind == A2CU_LIMITER || kind == A2CU_DCBLOCK) &&
        (nin != nout || nin < 1))
        return fail(A2CU_EINVAL, "a2cu_unit_alloc: unit needs matching i/o%s");
    if (kind == A2CU_FBDELAY && nin < 1) return fail(A2CU_EINVAL, "a2cu_unit_alloc: fbdelay needs an input%s");
    cudaSetDevice(e->device);
    int *fbd = nullptr;
    if (kind == A2CU_FBDELAY) {
        const size_t bytes = (size_t)2 * kFbdSize * sizeof(int);
        if (!e->fbd_free.empty()) { fbd = e->fbd_free.back(); e->fbd_free.pop_back(); }
        else if (cudaMalloc(&fbd, bytes) != cudaSuccess)
            return fail(A2CU_ENOMEM, "cudaMalloc fbdelay lines: %s", cudaGetErrorString(cudaGetLastError()));
        CK(cudaMemsetAsync(fbd, 0, bytes, e->stream));
    }
    int id;
    if (!e->gunit_free.empty()) { id = e->gunit_free.back(); e->gunit_free.pop_back(); }
    else {
        if ((int)e->gunits.size() + 1 > e->ustate_cap) {
            int ncap = std::max(256, e->ustate_cap * 2);
            int *n = nullptr;
            CK(cudaStreamSynchronize(e->stream));
            CK(cudaMalloc(&n, (size_t)ncap * kUnitWords * sizeof(int)));
            CK(memset_done(n, 0, (size_t)ncap * kUnitWords * sizeof(int)));
            if (e->d_ustate) {
                CK(memcpy_done(n, e->d_ustate, (size_t)e->ustate_cap * kUnitWords * sizeof(int),
                              cudaMemcpyDeviceToDevice));
                cudaFree(e->d_ustate);
            }
            e->d_ustate = n; e->ustate_cap = ncap;
        }
        id = (int)e->gunits.size();
        e->gunits.emplace_back();
    }
    a2cu_engine::GUnit &g = e->gunits[id];
    g.kind = kind; g.nin = nin; g.nout = nout; g.live = true; g.has_osc = false; g.fbd = fbd;
    return id;
}

int a2cu_unit_free(a2cu_engine *e, int unit) {
    if (!e || unit < 0 || unit >= (int)e->gunits.size() || !e->gunits[unit].live) return A2CU_EINVAL;
    e->gunits[unit].live = false;
    e->gunit_deferred.push_back(unit);      // still referenced by commands of this block
    return A2CU_OK;
}

static a2cu_engine::GUnit *get_gunit(a2cu_engine *e, int unit) {
    if (!e || unit < 0 || unit >= (int)e->gunits.size() || !e->gunits[unit].live) return nullptr;
    return &e->gunits[unit];
}

static void gunit_cmd(a2cu_engine *e, int op, int unit, const a2cu_engine::GUnit &g, int reg, int value, int start,
                      int dur) {
    BusCmd c;
    memset(&c, 0, sizeof(c));
    c.op = op; c.pm = unit; c.kind = g.kind; c.nin = g.nin; c.nout = g.nout;
    c.reg = reg; c.value = value; c.start = start; c.dur = dur;
    push_cmd(e, c);
}

int a2cu_block_unit_init(a2cu_engine *e, int unit, int transpose, unsigned substart) {
    a2cu_engine::GUnit *g = get_gunit(e, unit);
    if (!g) return fail(A2CU_EINVAL, "a2cu_block_unit_init: bad unit%s");
    int arg = init_arg(e, g->kind, transpose);
    if (g->kind == A2CU_FBDELAY) {
        // fbdelay.c:176-208: cleared delay lines, default registers
        const uint64_t ptr = (uint64_t)(uintptr_t)g->fbd;
        gunit_cmd(e, BUS_U_INIT, unit, *g, 0, (int)(uint32_t)ptr, 0, (int)(uint32_t)(ptr >> 32));
        static const int defaults[7] = {400 << 16, 280 << 16, 320 << 16, 65536, 16384, 32768, 32768};
        for (int r = 0; r < 7; ++r) {
            int r2 = a2cu_block_unit_write(e, unit, r, defaults[r], 0, 0, 0);
            if (r2) return r2;
        }
        return A2CU_OK;
    }
    gunit_cmd(e, BUS_U_INIT, unit, *g, 0, arg, (int)(substart & 0xff), 0);
    if (g->kind == A2CU_DCBLOCK)        // dcblock.c:127-128: default cutoff -5 (8.18 Hz), incl. transpose
        return a2cu_block_unit_write(e, unit, 0, (int)((unsigned)-5 << 16), transpose, 0, 0);
    if (g->kind == A2CU_FILTER12)
        gunit_cmd(e, BUS_U_WRITE, unit, *g, 5, tables().f12_coeff((int)((unsigned)arg << 8), e->samplerate), 0, 0);
    if (g->kind == A2CU_WTOSC) {
        // host twin of the control-rate part from the start: yields the noise
        // draw count of every Process() call (wtosc.c:135-144)
        g->osc.init(e, arg, substart & 0xff);
        g->has_osc = true;
    }
    return A2CU_OK;
}

int a2cu_block_unit_write(a2cu_engine *e, int unit, int reg, int32_t value, int transpose, unsigned start,
                          uint32_t dur) {
    a2cu_engine::GUnit *g = get_gunit(e, unit);
    if (!g) return fail(A2CU_EINVAL, "a2cu_block_unit_write: bad unit%s");
    Cooked c[2];
    int n = cook(e, g->kind, reg, value, (int)(start & 0xff), dur, transpose, c);
    if (n < 0) return n;
    if (g->has_osc) g->osc.write(e, c[0].reg, c[0].value, (int)(start & 0xff), (int)c[0].dur);
    for (int i = 0; i < n; ++i) gunit_cmd(e, BUS_U_WRITE, unit, *g, c[i].reg, c[i].value, (int)(start & 0xff), (int)c[i].dur);
    return A2CU_OK;
}

int a2cu_block_unit_proc(a2cu_engine *e, int unit, int add, int wireout, int scratch_bus, int out_bus, unsigned frame,
                         unsigned frames) {
    a2cu_engine::GUnit *g = get_gunit(e, unit);
    if (!g || scratch_bus < 0 || scratch_bus >= e->nbbus || frames < 1 || frame + frames > (unsigned)kMaxFrag ||
        (wireout && (out_bus < 0 || out_bus >= e->nbbus)))
        return fail(A2CU_EINVAL, "a2cu_block_unit_proc: bad args%s");
    if (g->has_osc) {
        bool is_noise;
        int draws = g->osc.segment(e, (int)frames, &is_noise);
        if (is_noise) {
            gunit_cmd(e, BUS_U_SEED, unit, *g, 0, (int)*e->noise_ptr, 0, 0);
            lcg_advance(e->noise_ptr, draws);
        }
    }
    BusCmd c;
    memset(&c, 0, sizeof(c));
    c.op = BUS_U_RUN; c.pm = unit; c.kind = g->kind; c.nin = g->nin; c.nout = g->nout;
    c.add = (add ? 1 : 0) | (wireout ? 2 : 0);
    c.in_bus = scratch_bus; c.out_bus = wireout ? out_bus : scratch_bus;
    c.frame = (int)frame; c.frames = (int)frames;
    push_cmd(e, c);
    return A2CU_OK;
}

int a2cu_block_bus_add(a2cu_engine *e, int src_bus, int dst_bus, unsigned frame, unsigned frames) {
    if (!e || src_bus < 0 || src_bus >= e->nbbus || dst_bus < 0 || dst_bus >= e->nbbus || frames < 1 ||
        frame + frames > (unsigned)kMaxFrag)
        return fail(A2CU_EINVAL, "a2cu_block_bus_add: bad args%s");
    BusCmd c;
    memset(&c, 0, sizeof(c));
    c.op = BUS_ADD; c.in_bus = src_bus; c.out_bus = dst_bus; c.frame = (int)frame; c.frames = (int)frames;
    push_cmd(e, c);
    return A2CU_OK;
}

static int block_flush_impl(a2cu_engine *e);
int a2cu_block_flush(a2cu_engine *e) {
    double t0 = now_us();
    int r = block_flush_impl(e);
    if (e) { e->bs.flush_us += now_us() - t0; ++e->bs.flushes; }
    return r;
}
static int block_flush_impl(a2cu_engine *e) {
    if (!e) return A2CU_EINVAL;
    cudaSetDevice(e->device);
    if (e->env_dry) {       // measurement aid (profiles/cfg5_k2trance.py --breakdown): drop what was recorded
        for (Bank *b : e->banks) { b->bev.clear(); b->bruns.clear(); b->cur_slot = -1; b->dup_runs = false; ++b->flush_id; }
        e->buscmds.clear(); e->runs.clear(); e->cur_run = -1; ++e->flush_serial;
        return A2CU_OK;
    }
    {
        int wr = upload_waves(e);           // waves first seen in this block
        if (wr) return wr;
    }
    for (Bank *b : e->banks) {
        if (!b->dynamic || b->bruns.empty()) continue;
        // One thread per voice: merge the runs of a slot (a voice is visited
        // once per segment of its parent) keeping record order.
        if (b->dup_runs) {
            std::vector<VoiceRun> sorted(b->bruns);
            std::stable_sort(sorted.begin(), sorted.end(),
                             [](const VoiceRun &a, const VoiceRun &c) { return a.slot < c.slot; });
            std::vector<uint4> ev;
            std::vector<VoiceRun> runs;
            ev.reserve(b->bev.size());
            for (const VoiceRun &r : sorted) {
                if (runs.empty() || runs.back().slot != r.slot) {
                    VoiceRun n;
                    n.slot = r.slot; n.ev_begin = (unsigned)ev.size(); n.ev_count = 0;
                    runs.push_back(n);
                }
                ev.insert(ev.end(), b->bev.begin() + r.ev_begin, b->bev.begin() + r.ev_begin + r.ev_count);
                runs.back().ev_count += r.ev_count;
            }
            b->bev.swap(ev);
            b->bruns.swap(runs);
        }
        b->dup_runs = false;
        ++b->flush_id;
        size_t nev = b->bev.size(), nr = b->bruns.size();
        if (nev > b->ev_cap) {
            CK(cudaStreamSynchronize(e->stream));
            if (b->d_ev) cudaFree(b->d_ev);
            b->ev_cap = nev * 2;
            CK(cudaMalloc(&b->d_ev, b->ev_cap * sizeof(uint4)));
        }
        if (nr > b->runs_cap) {
            CK(cudaStreamSynchronize(e->stream));
            if (b->d_runs) cudaFree(b->d_runs);
            b->runs_cap = nr * 2;
            CK(cudaMalloc(&b->d_runs, b->runs_cap * sizeof(VoiceRun)));
        }
        CK(cudaMemcpyAsync(b->d_ev, b->bev.data(), nev * sizeof(uint4), cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemcpyAsync(b->d_runs, b->bruns.data(), nr * sizeof(VoiceRun), cudaMemcpyHostToDevice, e->stream));
        e->h2d_bytes += nev * sizeof(uint4) + nr * sizeof(VoiceRun);
        RenderParams P;
        memset(&P, 0, sizeof(P));
        P.state = b->d_state; P.stride = b->stride; P.nvoices = (int)nr;
        P.acc = e->d_bacc; P.W = kMaxFrag; P.buffer = kMaxFrag;
        P.waves = e->d_waves; P.pool = e->d_pool; P.cpool = e->d_cpool; P.ptab = e->d_ptab; P.fmsine = e->d_fmsine;
        P.f12tab = e->d_f12tab;
        P.samplerate = e->samplerate;
        P.ev = b->d_ev; P.runs = b->d_runs; P.explicit_ = 1;
        int grid = ((int)nr + kThreads - 1) / kThreads;
        if (b->generic) render_generic<<<grid, kThreads, 0, e->stream>>>(P, b->gchain, b->d_scratch);
        else b->k.fn<<<grid, kThreads, 0, e->stream>>>(P);
        ++e->launches;
        b->bev.clear(); b->bruns.clear(); b->cur_slot = -1;
    }
    if (!e->buscmds.empty()) {
        // Order runs by nest level, deepest first (stable), and lay their
        // commands out contiguously; one bus_level launch per level.
        const size_t n = e->buscmds.size(), nr = e->runs.size();
        int maxlevel = 0;
        for (const auto &r : e->runs) maxlevel = std::max(maxlevel, r.level);
        std::vector<unsigned> level_runs(maxlevel + 2, 0);
        for (const auto &r : e->runs) if (r.count) ++level_runs[r.level];
        // run order: level maxlevel first
        std::vector<unsigned> level_first(maxlevel + 2, 0);
        unsigned pos = 0;
        for (int l = maxlevel; l >= 0; --l) { level_first[l] = pos; pos += level_runs[l]; }
        const unsigned nlive = pos;
        e->sorted_runs.assign(nlive, BusRun{0, 0});
        std::vector<unsigned> run_pos(nr, 0xffffffffu), fill(level_first);
        for (size_t i = 0; i < nr; ++i)
            if (e->runs[i].count) run_pos[i] = fill[e->runs[i].level]++;
        for (size_t i = 0; i < nr; ++i)
            if (e->runs[i].count) e->sorted_runs[run_pos[i]].count = e->runs[i].count;
        unsigned cpos = 0;
        for (unsigned i = 0; i < nlive; ++i) { e->sorted_runs[i].begin = cpos; cpos += e->sorted_runs[i].count; }
        e->sorted_cmds.resize(n);
        std::vector<unsigned> wr(nlive);
        for (unsigned i = 0; i < nlive; ++i) wr[i] = e->sorted_runs[i].begin;
        for (size_t i = 0; i < n; ++i) e->sorted_cmds[wr[run_pos[e->buscmds[i].run]]++] = e->buscmds[i];
        if (n > e->buscmds_cap) {
            CK(cudaStreamSynchronize(e->stream));
            if (e->d_buscmds) cudaFree(e->d_buscmds);
            e->buscmds_cap = n * 2;
            CK(cudaMalloc(&e->d_buscmds, e->buscmds_cap * sizeof(BusCmd)));
        }
        if (nlive > e->runs_cap) {
            CK(cudaStreamSynchronize(e->stream));
            if (e->d_runs) cudaFree(e->d_runs);
            e->runs_cap = (size_t)nlive * 2;
            CK(cudaMalloc(&e->d_runs, e->runs_cap * sizeof(BusRun)));
        }
        CK(cudaMemcpyAsync(e->d_buscmds, e->sorted_cmds.data(), n * sizeof(BusCmd), cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemcpyAsync(e->d_runs, e->sorted_runs.data(), nlive * sizeof(BusRun), cudaMemcpyHostToDevice,
                           e->stream));
        e->h2d_bytes += n * sizeof(BusCmd) + nlive * sizeof(BusRun);
        BusVmParams BP;
        memset(&BP, 0, sizeof(BP));
        BP.cmds = e->d_buscmds; BP.acc = e->d_bacc; BP.pmstate = e->d_pmstate;
        BP.ustate = e->d_ustate;
        BP.ctx.waves = e->d_waves; BP.ctx.pool = e->d_pool; BP.ctx.cpool = e->d_cpool; BP.ctx.ptab = e->d_ptab;
        BP.ctx.fmsine = e->d_fmsine; BP.ctx.f12tab = e->d_f12tab; BP.ctx.samplerate = e->samplerate;
        for (int l = maxlevel; l >= 0; --l) {
            if (!level_runs[l]) continue;
            BP.runs = e->d_runs + level_first[l];
            bus_level<<<level_runs[l], kMaxFrag, 0, e->stream>>>(BP);
            ++e->launches;
        }
        e->buscmds.clear(); e->runs.clear(); e->cur_run = -1; ++e->flush_serial;
    }
    CK(cudaGetLastError());
    return A2CU_OK;
}

static int ensure_xfer(a2cu_engine *e) {
    if (!e->h_xfer) CK(cudaMallocHost(&e->h_xfer, kMaxFrag * 2 * sizeof(int32_t)));
    return A2CU_OK;
}

int a2cu_block_upload(a2cu_engine *e, int bus, int nch, unsigned frame, unsigned frames, const int32_t *const *src) {
    if (!e || bus < 0 || bus >= e->nbbus || frames < 1 || frame + frames > (unsigned)kMaxFrag || nch < 1 || nch > 2)
        return fail(A2CU_EINVAL, "a2cu_block_upload: bad args%s");
    cudaSetDevice(e->device);
    int r = a2cu_block_flush(e);            // keep device order == host walk order
    if (r) return r;
    r = ensure_xfer(e);
    if (r) return r;
    CK(cudaStreamSynchronize(e->stream));   // h_xfer reuse
    for (unsigned i = 0; i < frames; ++i) {
        e->h_xfer[i * 2] = src[0][frame + i];
        e->h_xfer[i * 2 + 1] = nch > 1 ? src[1][frame + i] : 0;
    }
    CK(cudaMemcpyAsync(e->d_bacc + ((size_t)bus * kMaxFrag + frame) * 2, e->h_xfer, frames * 2 * sizeof(int32_t),
                       cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->h2d_bytes += frames * 2 * sizeof(int32_t);
    return A2CU_OK;
}

int a2cu_block_download(a2cu_engine *e, int bus, int nch, unsigned frame, unsigned frames, int32_t *const *dst,
                        int add) {
    if (!e || bus < 0 || bus >= e->nbbus || frames < 1 || frame + frames > (unsigned)kMaxFrag || nch < 1 || nch > 2)
        return fail(A2CU_EINVAL, "a2cu_block_download: bad args%s");
    cudaSetDevice(e->device);
    int r = a2cu_block_flush(e);
    if (r) return r;
    r = ensure_xfer(e);
    if (r) return r;
    if (e->env_dry) {
        if (!add) for (int c = 0; c < nch; ++c) memset(dst[c] + frame, 0, frames * sizeof(int32_t));
        return A2CU_OK;
    }
    double t0 = now_us();
    CK(cudaMemcpyAsync(e->h_xfer, e->d_bacc + ((size_t)bus * kMaxFrag + frame) * 2, frames * 2 * sizeof(int32_t),
                       cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->bs.sync_us += now_us() - t0;
    ++e->bs.downloads;
    e->d2h_bytes += frames * 2 * sizeof(int32_t);
    for (int c = 0; c < nch; ++c)
        for (unsigned i = 0; i < frames; ++i) {
            if (add) dst[c][frame + i] += e->h_xfer[i * 2 + c];
            else dst[c][frame + i] = e->h_xfer[i * 2 + c];
        }
    return A2CU_OK;
}

}  // extern "C"
