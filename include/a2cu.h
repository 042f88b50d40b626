/*
 * a2cu.h - C ABI of the B200 voice-render engine ("bank mode").
 *
 * This is the device engine underneath the drop-in unit plug-in
 * (include/a2cu_units.h) and the entry point for hosts that batch voices
 * themselves.  Plain C: opaque handle, pointers and sizes, no torch or CUDA
 * types in the signatures (a CUDA stream travels as void *).
 *
 * Relation to the reference (paths relative to the Audiality 2 tree):
 *
 *   a2cu_open / a2cu_close     a2_Open / a2_Close for the render path only:
 *                              sample rate, channel count, basepitch
 *                              (src/audiality2.c:398-399, 406-511)
 *   a2cu_wave_*                read side of the wave store: a2_InitWaves
 *                              (src/waves.c:629-708), pad + mip preparation
 *                              (src/waves.c:59-151), a2_GetWave (:759-772)
 *   a2cu_bank_new              a2_PopulateVoice / a2_AddUnit + each unit's
 *                              Initialize() (src/core.c:163-420) for N voices
 *                              of one struct at once
 *   a2cu_bank_write[_all]      A2_write_cb of the unit's A2_crdesc
 *                              (include/a2_units.h:115), i.e. what
 *                              a2_VoiceControl delivers (src/core.c:143-149)
 *   a2cu_bank_wake             a VM wake-up that writes nothing but still
 *                              splits the Process() segment (core.c:1847-1880)
 *   a2cu_group_new/_write      a2_NewGroup's a2_groupdriver voice:
 *                              inline; panmix; xinsert (audiality2.c:294-304)
 *   a2cu_root_write            the root driver's panmix (audiality2.c:268-292)
 *   a2cu_run                   a2_Run(): a2_AudioCallback's fragment loop,
 *                              a2_ProcessVoices, every unit's Process() and
 *                              the bus mix-down (src/core.c:1847-2011); output
 *                              is what the driver's int32 8:24 buffers hold
 *                              (include/a2_drivers.h:301), interleaved
 *
 * Values are 16:16 fixed point exactly as the VM registers hold them; times
 * are 24:8 fixed point sample frames since a2cu_open (A2_timestamp units).
 * Every function returns 0 (A2CU_OK) or a negative a2cu_error unless stated.
 * There is NO CPU fallback: without a usable CUDA device a2cu_open fails.
 */
#ifndef A2CU_H
#define A2CU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct a2cu_engine a2cu_engine;

typedef enum a2cu_error {
	A2CU_OK = 0,
	A2CU_ENODEVICE = -1,	/* no CUDA device / wrong architecture */
	A2CU_ECUDA = -2,	/* CUDA runtime error, see a2cu_last_error() */
	A2CU_EINVAL = -3,	/* bad argument */
	A2CU_ENOTIMPL = -4,	/* voice structure has no kernel (A2_NOTIMPLEMENTED) */
	A2CU_ENOMEM = -5,
	A2CU_ELATE = -6		/* event time already rendered */
} a2cu_error;

/* Unit kinds (a2_core_units[], src/audiality2.c:183-207) */
enum {
	A2CU_WTOSC = 1, A2CU_PANMIX = 2, A2CU_FILTER12 = 3, A2CU_WAVESHAPER = 4,
	A2CU_FBDELAY = 5,	/* units/fbdelay.c; generic unit only */
	A2CU_LIMITER = 6,	/* units/limiter.c; generic unit only */
	A2CU_DCBLOCK = 7,	/* units/dcblock.c; generic unit only */
	A2CU_DC = 8,		/* units/dc.c; generic unit only */
	A2CU_FM1 = 16, A2CU_FM2, A2CU_FM3, A2CU_FM4,
	A2CU_FM3P, A2CU_FM4P, A2CU_FM2R, A2CU_FM4R
};

/* Wave types and flags (include/a2_waves.h:78-84, :108) */
enum { A2CU_WOFF = 0, A2CU_WNOISE = 1, A2CU_WWAVE = 2, A2CU_WMIPWAVE = 3 };
#define A2CU_LOOPED	0x100
#define A2CU_MIPLEVELS	10

/*
 * One unit of a voice structure after the compiler's autowiring
 * (src/compiler.c:3036-3138) and instantiation (src/core.c:163-243).
 */
typedef struct a2cu_unitspec {
	int32_t kind;
	int32_t ninputs;
	int32_t noutputs;
	int32_t add;		/* A2_PROCADD */
	int32_t wireout;	/* A2_IO_WIREOUT: outputs are the voice's bus */
} a2cu_unitspec;

/* ---- engine ---------------------------------------------------------- */

/* device: CUDA ordinal. channels: 1 or 2. Returns NULL on failure. */
a2cu_engine *a2cu_open(int device, int samplerate, int channels);
void a2cu_close(a2cu_engine *e);
const char *a2cu_last_error(void);

/* Use an existing CUDA stream (cudaStream_t passed as void *); default 0. */
int a2cu_set_stream(a2cu_engine *e, void *cuda_stream);

int a2cu_basepitch(const a2cu_engine *e);		/* audiality2.c:398 */
uint32_t a2cu_msdur(const a2cu_engine *e);		/* audiality2.c:499 */
uint64_t a2cu_now(const a2cu_engine *e);		/* 24:8 frames rendered */

/*
 * The reference's root voice re-arms a wake-up every 1000000 time units
 * (src/core.c:1191-1216), which splits every voice's segments there.  On by
 * default so that output equals the reference's offline render; 0 disables.
 */
int a2cu_set_root_wake_period(a2cu_engine *e, uint32_t period_24_8);

/* Seed of the shared noise LCG (A2_PNOISESEED, src/properties.c:298). */
int a2cu_set_noiseseed(a2cu_engine *e, uint32_t seed);
/*
 * Drop-in mode: use the host's own LCG word (st->noisestate, src/internals.h:
 * 682) so the VM's RAND instructions and the noise oscillators keep sharing one
 * sequence (src/core.c:1401-1409, src/units/wtosc.c:135-144).
 */
int a2cu_set_noise_state_ptr(a2cu_engine *e, uint32_t *state);

/* ---- waves ----------------------------------------------------------- */

/* Builtin wave by its A2S name ("sine", "saw", "pulse25", ...). Returns id. */
int a2cu_wave_builtin(a2cu_engine *e, const char *name);

/* Upload raw int16 samples; pads and mip levels are prepared here. Returns id. */
int a2cu_wave_upload(a2cu_engine *e, int type, unsigned period, unsigned flags,
		const int16_t *data, unsigned length);

/*
 * Upload a wave the host already prepared (A2_wave, include/a2_waves.h:88-103):
 * data[l] points at the FIRST PAD sample of level l (A2_wave_wave.data[l]),
 * size[l] excludes the pads. Used by the drop-in plug-in. Returns id.
 */
int a2cu_wave_upload_prepared(a2cu_engine *e, int type, unsigned period,
		unsigned flags, const int16_t *const *data, const unsigned *size);

/* Mark a wave unloaded (size[0] = 0, src/waves.c:718-724). */
int a2cu_wave_unload(a2cu_engine *e, int wave);

/* Copy level 'level' incl. pads to 'out' (capacity in samples); returns count. */
int a2cu_wave_read(a2cu_engine *e, int wave, int level, int16_t *out,
		unsigned capacity, unsigned *size);

/* ---- groups and banks -------------------------------------------------- */

/* New group bus with its own panmix, mixing into the root bus. Returns id. */
int a2cu_group_new(a2cu_engine *e);

/* 1 if a kernel exists for this voice structure. */
int a2cu_chain_supported(const a2cu_unitspec *chain, int nunits);

/*
 * Create a bank: nvoices voices of one structure, started at the current time
 * with sub-sample offset 'substart' (waketime & 0xff). transpose / group may
 * be NULL (0 / root bus) or arrays of nvoices entries. Returns bank id.
 */
int a2cu_bank_new(a2cu_engine *e, const a2cu_unitspec *chain, int nunits,
		int nvoices, const int32_t *transpose, const int32_t *group,
		unsigned substart);

/* Stop a voice at 'when' (a2_VoiceFree at a segment start, core.c:1891). */
int a2cu_bank_kill(a2cu_engine *e, int bank, int voice, uint64_t when);

/* One control write to one voice. 'when' >= a2cu_now(). */
int a2cu_bank_write(a2cu_engine *e, int bank, int voice, int unit, int reg,
		int32_t value, uint64_t when, uint32_t dur);

/*
 * The same register of every voice of the bank: values[i * stride] for voice
 * i (stride 0 broadcasts values[0]).
 */
int a2cu_bank_write_all(a2cu_engine *e, int bank, int unit, int reg,
		const int32_t *values, int stride, uint64_t when, uint32_t dur);

/*
 * Pause (0) / resume (1) a bank. A paused bank is skipped by a2cu_run*: its
 * voices keep their state and its pending writes apply at the start of the
 * first window it runs in again. One engine can so serve many banks
 * round-robin (bench.py uses this to keep per-step state colder than L2).
 */
int a2cu_bank_enable(a2cu_engine *e, int bank, int enabled);

/* Bare wake-up (segment split) of one voice, or of all with voice < 0. */
int a2cu_bank_wake(a2cu_engine *e, int bank, int voice, uint64_t when);

/* reg: 0 = vol, 1 = pan (src/units/panmix.c:29-33); reg < 0: bare wake-up */
int a2cu_group_write(a2cu_engine *e, int group, int reg, int32_t value,
		uint64_t when, uint32_t dur);
int a2cu_root_write(a2cu_engine *e, int reg, int32_t value, uint64_t when,
		uint32_t dur);

/* ---- rendering ----------------------------------------------------------- */

/*
 * Render 'frames' frames as driver buffers of 'buffer' frames (fragments of at
 * most 64 inside each buffer, src/core.c:1964-1973) and copy the interleaved
 * int32 8:24 master output [frame][channel] to HOST memory 'out'. Blocks
 * until the copy is complete. 'out' may be NULL (render only).
 */
int a2cu_run(a2cu_engine *e, unsigned frames, unsigned buffer, int32_t *out);

/*
 * Same, but leaves the result in DEVICE memory 'dev_out' (or in the engine's
 * own buffer when NULL; see a2cu_master_devptr) and does not synchronise: work
 * is queued on the engine's stream.
 */
int a2cu_run_async(a2cu_engine *e, unsigned frames, unsigned buffer,
		int32_t *dev_out);
int32_t *a2cu_master_devptr(a2cu_engine *e);
/*
 * Sample format of the master block that a2cu_run / a2cu_collect / dev_out
 * deliver, converted inside the root stage (no extra pass): what the
 * reference's drivers and wave writer do with the int32 8:24 buffers at the
 * edge of the engine.
 *   0  int32 8:24 (include/a2_drivers.h:301) - default
 *   1  float32 = v * (1 / 8388608)           (src/drivers/sdldrv.c:55-65)
 *   2  int16   = v >> 8                      (a2_RenderWave -> a2_WaveWrite with
 *                                             A2_I24, src/waves.c:174-176)
 * The caller's buffer holds frames * channels samples of that size.
 */
int a2cu_set_output_format(a2cu_engine *e, int format);
int a2cu_sync(a2cu_engine *e);

/*
 * Pipelined form of a2cu_run for offline rendering (a2_Render, src/render.c:
 * 34-127, renders buffer after buffer without waiting for a consumer):
 * a2cu_submit queues one window - event staging, H2D, kernels, D2H of the
 * master block into a pinned result slot - and returns a ticket (>= 0) without
 * waiting; a2cu_collect blocks until THAT window is done and copies its
 * int32 8:24 output to 'out'.  Up to 4 windows may be in flight; the host
 * prepares window i+1 (a2cu_bank_write*) while the device renders window i.
 */
int a2cu_submit(a2cu_engine *e, unsigned frames, unsigned buffer);
/* Output stays in DEVICE memory 'dev_out' (multi-GPU: the raw root bus that is
 * reduced next); a2cu_collect(e, ticket, NULL) waits for the window's kernels. */
int a2cu_submit_dev(a2cu_engine *e, unsigned frames, unsigned buffer,
		int32_t *dev_out);
int a2cu_collect(a2cu_engine *e, int ticket, int32_t *out);

/*
 * Multi-GPU cut: with post_root_stage 0 the engine stops
 * before the root panmix and outputs the raw 2-channel root bus, so partial
 * buses of several engines can be summed (integer add, order-free) before ONE
 * engine applies the truncating root stage with a2cu_apply_root_stage().
 */
int a2cu_set_post_root_stage(a2cu_engine *e, int enabled);
int a2cu_apply_root_stage(a2cu_engine *e, const int32_t *dev_rootbus,
		int32_t *dev_master, unsigned frames, unsigned buffer,
		uint64_t start_time);

/*
 * Sharded render with the exchange INSIDE the render kernel (no NCCL call, no
 * extra launch): every rank creates a "symmetric" buffer, the ranks swap its
 * handle (64 bytes; over torch.distributed, MPI, a pipe ... - plumbing the
 * caller owns) and map each other's buffers (CUDA IPC between processes, peer
 * access inside one process). From then on each a2cu_run / a2cu_submit window
 * ends like this on every rank: the CTA holding the finished root bus stores
 * it into all peers' buffers over NVLink, releases a flag, waits for the
 * world's flags, sums the rows and runs the root stage - all ranks obtain the
 * identical master block (integer sum: bit-identical to one engine rendering
 * all voices). All ranks must render the same sequence of windows.
 *   rank, world   0 <= rank < world <= 8
 *   max_frames    longest window that will be rendered
 *   timeout_ms    bound of the in-kernel wait for peers (0: 2000); on expiry
 *                 the window's a2cu_collect / a2cu_sync returns A2CU_ECUDA
 *   handle_out    64 bytes (cudaIpcMemHandle_t) or NULL
 */
int a2cu_xchg_create(a2cu_engine *e, int rank, int world, unsigned max_frames,
		unsigned timeout_ms, void *handle_out);
/* handles: world x 64 bytes in rank order (the own entry is ignored). */
int a2cu_xchg_connect_ipc(a2cu_engine *e, const void *handles);
/* Engines of ONE process (any devices): peers[r] = the engine of rank r. */
int a2cu_xchg_connect_local(a2cu_engine *e, a2cu_engine *const *peers);
/* Turn the exchange off / on again without unmapping (A/B measurements). */
int a2cu_xchg_enable(a2cu_engine *e, int enabled);
int a2cu_xchg_close(a2cu_engine *e);

/* Kernel launches issued by this engine so far (bench's gpu_launches). */
uint64_t a2cu_launch_count(const a2cu_engine *e);
/*
 * Wavetable voice structures have a second, warp-specialised kernel
 * (render_split, csrc/a2cu_split.cuh) that the engine picks per launch when
 * the window qualifies; a2cu_set_split(e, 0) forces the one-thread-per-voice
 * kernel (same results, used for A/B tests). Also: env A2CU_NO_SPLIT.
 */
int a2cu_set_split(a2cu_engine *e, int enabled);
uint64_t a2cu_split_launch_count(const a2cu_engine *e);
/*
 * Profiling aid: per-role busy cycles of render_split summed over CTAs since
 * the last call: out[0] control, [1] filter recurrence (computing), [2]
 * oscillator stage and [3] panmix/bus stage of one helper warp per voice set,
 * [4] filter recurrence warp waiting for its input, [5] fragments rendered,
 * [6] kernel entry -> pipeline start, [7] pipeline + state store (per CTA).
 * enable != 0 (re)arms the counters, 0 turns them off. out may be NULL.
 */
int a2cu_split_profile(a2cu_engine *e, int enable, uint64_t out[8]);
/*
 * Timeline of CTA 0 / voice set 0 of the last render_split launch while the
 * profile is armed: out[(role * 64 + fragment) * 2 + end] = cycles since the
 * pipeline start; roles 0 control, 1 filter recurrence, 2 / 3 oscillator /
 * panmix stage of helper 0, 4 / 5 of the last helper. out holds 768 words.
 */
int a2cu_split_trace(a2cu_engine *e, uint64_t *out);
/* Re-arm the grid-wide wall-clock marks (role 4, fragments 56..59: first CTA entry, last pipeline
 * end, last CTA done, end of the fused tail; globaltimer ns) before the launch to be looked at. */
int a2cu_split_trace_reset(a2cu_engine *e);
/* Name of the render kernel a bank uses (for profiles/). */
const char *a2cu_bank_kernel_name(a2cu_engine *e, int bank);
/* Bytes of per-voice state a bank keeps in HBM (roofline arithmetic). */
int a2cu_bank_state_bytes(a2cu_engine *e, int bank);
/* CUDA-event time (ms) of the last a2cu_run*'s render kernels only. */
float a2cu_last_render_ms(a2cu_engine *e);
/* Same for the bus stage (mix_buses) of the last run. */
float a2cu_last_mix_ms(a2cu_engine *e);
/* Bytes this engine copied host->device / device->host so far. */
uint64_t a2cu_h2d_bytes(const a2cu_engine *e);
uint64_t a2cu_d2h_bytes(const a2cu_engine *e);
/* Turn the event timing above on/off (adds three event records per run). */
int a2cu_set_timing(a2cu_engine *e, int enabled);
/*
 * Self-test hook: f12_pitch2coeff (src/units/filter12.c:65-72) as the kernels
 * evaluate it, for n cutoff ramper values (8:24 fixed point, i.e. what
 * A2_filter12.cutoff.value holds). The device looks the coefficient up in a
 * table the host built with its own libm for every possible argument, so the
 * result must equal the reference's for ALL inputs (tests sweep the domain).
 */
int a2cu_debug_f12_coeff(a2cu_engine *e, const int32_t *cutoff_values, int n,
		int32_t *out);

/* ---- drop-in ("block") mode ------------------------------------------------
 *
 * What the unit plug-in (include/a2cu_units.h, plugin/a2cu_units.c) records
 * while the UNMODIFIED reference host walks its voice tree, one fragment of at
 * most 64 frames at a time (a2_AudioCallback, src/core.c:1964-1973):
 *
 *   Initialize()  of a voice's units  -> a2cu_pool_alloc + a2cu_block_init
 *   A2_write_cb                       -> a2cu_block_write   (leaf voices)
 *                                        a2cu_block_pm_write (bus-level panmix)
 *   Process(u, offset, frames)        -> a2cu_block_proc     (one per voice and
 *                                        segment, src/core.c:1875-1876)
 *                                        a2cu_block_pm_proc  (bus-level panmix)
 *   inline's Process (src/core.c:1763-1776) -> a2cu_block_bus: a fresh device
 *                                        bus for the sub-tree being walked
 *   a unit whose successor is a host unit -> a2cu_block_download (flushes the
 *                                        recorded work, then copies the bus)
 *   a unit whose input was written by a host unit -> a2cu_block_upload
 *   Deinitialize()                    -> a2cu_pool_free / a2cu_pm_free
 *
 * 'frame' arguments are fragment-relative (0..63).
 */
int a2cu_pool_open(a2cu_engine *e, const a2cu_unitspec *chain, int nunits);
int a2cu_pool_alloc(a2cu_engine *e, int pool);
int a2cu_pool_free(a2cu_engine *e, int pool, int slot);
int a2cu_block_begin(a2cu_engine *e);
int a2cu_block_bus(a2cu_engine *e);
int a2cu_block_init(a2cu_engine *e, int pool, int slot, int unit, int transpose,
		unsigned frame, unsigned substart);
int a2cu_block_write(a2cu_engine *e, int pool, int slot, int unit, int reg,
		int32_t value, int transpose, unsigned frame, unsigned start,
		uint32_t dur);
int a2cu_block_proc(a2cu_engine *e, int pool, int slot, unsigned frame,
		unsigned frames, int bus);
int a2cu_pm_alloc(a2cu_engine *e);
int a2cu_pm_free(a2cu_engine *e, int pm);
int a2cu_block_pm_write(a2cu_engine *e, int pm, int reg, int32_t value,
		unsigned start, uint32_t dur);
int a2cu_block_pm_proc(a2cu_engine *e, int pm, int nin, int nout, int add,
		int in_bus, int out_bus, unsigned frame, unsigned frames);
/*
 * Generic units: ONE replaced unit called on its own - bus-level filter12 /
 * waveshaper / wtosc after an `inline`, or our units inside a chain that also
 * holds host units ({wtosc; panmix 1 2; fbdelay 2 >}).  The voice's scratch
 * channels (st->scratch[nest], src/core.c:364-395) live in a device bus row;
 * a2cu_block_unit_proc is one Process(u, frame, frames) call (a2_units.h:176)
 * in the host's order, 'add' = A2_PROCADD, 'wireout' = outputs are the voice's
 * output bus 'out_bus' (A2_IO_WIREOUT, src/core.c:243-245).
 */
/*
 * Bus commands (a2cu_block_pm_*, a2cu_block_unit_*, a2cu_block_bus_add) belong
 * to the voice selected by the last a2cu_block_run(): 'level' is the voice's
 * nest level (A2_voice.nestlevel, src/internals.h:571), 'prev' the handle this
 * function returned for the same voice earlier (0 for none).  Voices of one
 * level run concurrently on the device, levels deepest first; commands of one
 * voice keep their order.  Every flush (a2cu_block_flush / _upload / _download /
 * _begin) ends all runs: select the voice again before its next command.
 */
uint64_t a2cu_block_run(a2cu_engine *e, int level, uint64_t prev);
int a2cu_unit_alloc(a2cu_engine *e, int kind, int ninputs, int noutputs);
int a2cu_unit_free(a2cu_engine *e, int unit);
int a2cu_block_unit_init(a2cu_engine *e, int unit, int transpose, unsigned substart);
int a2cu_block_unit_write(a2cu_engine *e, int unit, int reg, int32_t value,
		int transpose, unsigned start, uint32_t dur);
int a2cu_block_unit_proc(a2cu_engine *e, int unit, int add, int wireout,
		int scratch_bus, int out_bus, unsigned frame, unsigned frames);
/* dst bus += src bus over [frame, frame + frames) */
int a2cu_block_bus_add(a2cu_engine *e, int src_bus, int dst_bus, unsigned frame,
		unsigned frames);
int a2cu_block_flush(a2cu_engine *e);
int a2cu_block_upload(a2cu_engine *e, int bus, int nch, unsigned frame,
		unsigned frames, const int32_t *const *src);
int a2cu_block_download(a2cu_engine *e, int bus, int nch, unsigned frame,
		unsigned frames, int32_t *const *dst, int add);

#ifdef __cplusplus
}
#endif
#endif /* A2CU_H */
