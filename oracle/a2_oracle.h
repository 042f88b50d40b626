/*
 * a2_oracle.h - CPU restatement ("port") of Audiality 2's per-voice DSP path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.
 * The product (audiality2_b200/) never links, imports or executes it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this port bit for
 * bit against outputs of the reference itself (oracle/_ref/libaudiality2.so,
 * built by oracle/Makefile from /root/reference) and against the golden
 * fixtures under tests/golden/ that were generated from that build with
 * tests/golden/make_golden.py.  The reference ships no golden vectors of its
 * own (its tests are listen-tests).
 *
 * What is restated (reference file:line in a2_oracle.c next to each function):
 *   include/a2_dsp.h        noise LCG, Lerp, Hermite, ramper
 *   src/pitch.c:57-96       a2_P2I and its table
 *   src/waves.c:90-151, 629-708   pad/mip preparation, builtin waves
 *   src/units/wtosc.c       wavetable oscillator, noise, off
 *   src/units/panmix.c      1/2 -> 1/2 pan/mix
 *   src/units/filter12.c    12 dB state variable filter
 *   src/units/fm.c          fm1..fm4, fm3p, fm4p, fm2r, fm4r
 *   src/units/waveshaper.c  rational waveshaper
 *   src/core.c:1847-1896, 1749-1776   segment loop, add/replace bus semantics
 *
 * What is NOT restated: the A2S compiler and VM.  Callers describe what the
 * VM would have done as an explicit, time-ordered event list (control writes
 * and bare wake-ups); the event list of every golden test is paired with the
 * .a2s script it mirrors.
 */
#ifndef A2_ORACLE_H
#define A2_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A2O_MAXFRAG	64	/* include/audiality2.h.cmake:50 */
#define A2O_MIPLEVELS	10	/* include/a2_waves.h:33 */
#define A2O_WAVEPRE	1	/* include/a2_waves.h:61 */
#define A2O_WAVEPOST	132	/* include/a2_waves.h:64 (2 + 129 + 1) */
#define A2O_MAXUNITS	12

/* Unit kinds. The fm values encode operators and structure. */
enum {
	A2O_WTOSC = 1, A2O_PANMIX, A2O_FILTER12, A2O_WAVESHAPER,
	A2O_FM1 = 16, A2O_FM2, A2O_FM3, A2O_FM4,
	A2O_FM3P, A2O_FM4P, A2O_FM2R, A2O_FM4R
};

/* Wave types, include/a2_waves.h:78-84 */
enum { A2O_WOFF = 0, A2O_WNOISE, A2O_WWAVE, A2O_WMIPWAVE };
#define A2O_LOOPED	0x100	/* include/a2_waves.h:108 */

/* Control register indices, in the reference's A2_crdesc order */
enum { A2O_W_WAVE = 0, A2O_W_PITCH, A2O_W_AMP, A2O_W_PHASE };	/* wtosc.c:58-64 */
enum { A2O_PM_VOL = 0, A2O_PM_PAN };				/* panmix.c:29-33 */
enum { A2O_F_CUTOFF = 0, A2O_F_Q, A2O_F_LP, A2O_F_BP, A2O_F_HP };/* filter12.c:27-34 */
enum { A2O_WS_AMOUNT = 0 };					/* waveshaper.c:30-33 */
/* fm: 0 = phase, then (p, a, fb) per operator; fm.c:53-76 */
enum { A2O_FM_PHASE = 0, A2O_FM_P0, A2O_FM_A0, A2O_FM_FB0 };

/* One unit of a voice structure, after autowiring (compiler.c:3036-3138) */
typedef struct a2o_unitspec
{
	int	kind;
	int	ninputs;	/* scratch channels read */
	int	noutputs;	/* channels written */
	int	add;		/* A2_PROCADD */
	int	wireout;	/* outputs go to the voice's output bus */
} a2o_unitspec;

/* Event kinds for a2o_render() */
enum {
	A2O_EV_WRITE = 0,	/* control register write */
	A2O_EV_WAKE,		/* VM woke up and wrote nothing: splits only */
	A2O_EV_ROOTWRITE,	/* write to the root panmix (unit = 0) */
	A2O_EV_GROUPWRITE	/* write to a group's panmix; voice = group */
};

typedef struct a2o_event
{
	uint32_t	time;	/* 24:8 frames since start (VM waketime) */
	int32_t		kind;
	int32_t		voice;
	int32_t		unit;
	int32_t		reg;
	int32_t		value;	/* 16:16, as the VM register holds it */
	uint32_t	dur;	/* 24:8 ramp duration */
} a2o_event;

typedef struct a2o_engine a2o_engine;

a2o_engine *a2o_open(int samplerate, int channels);
void a2o_close(a2o_engine *e);
int a2o_basepitch(a2o_engine *e);
uint32_t a2o_msdur(a2o_engine *e);
void a2o_set_noiseseed(a2o_engine *e, uint32_t seed);

/* Waves. Return wave id >= 0 or -1. */
int a2o_builtin_wave(a2o_engine *e, const char *name);
int a2o_upload_wave(a2o_engine *e, int type, unsigned period, unsigned flags,
		const int16_t *data, unsigned length);
/* Access to prepared wave data (for cross-checking the product's tables) */
const int16_t *a2o_wave_data(a2o_engine *e, int wave, int level,
		unsigned *size);

/* Groups (a2_groupdriver-like: inline; panmix; add to root bus). */
int a2o_new_group(a2o_engine *e);
/* song-level chain { inline 0 *; fbdelay * *; panmix * > } (units/fbdelay.c): 7 registers, 16:16 */
int a2o_group_fbdelay(a2o_engine *e, int group, const int32_t *regs);

/*
 * Voices. 'transpose' is the voice's R_TRANSPOSE (16:16), 'substart' the
 * sub-sample start time (waketime & 0xff), 'group' -1 for the root bus.
 */
int a2o_new_voice(a2o_engine *e, const a2o_unitspec *chain, int nunits,
		int transpose, unsigned substart, int group);
void a2o_kill_voice(a2o_engine *e, int voice);

/* Low level: immediate control write / one Process() round for one voice */
void a2o_write(a2o_engine *e, int voice, int unit, int reg, int value,
		unsigned start, unsigned dur);

void a2o_write_all(a2o_engine *e, int first, int count, int unit, int reg,
		const int *values, int stride, unsigned start, unsigned dur);

/*
 * Render 'frames' frames in driver buffers of 'buffer' frames, fragments of
 * at most A2O_MAXFRAG, applying 'ev' (sorted by time; stable order within a
 * time stamp is preserved per voice).  'out' is interleaved int32 8:24,
 * [frame][channel].  Time continues from the previous call.
 */
void a2o_render(a2o_engine *e, const a2o_event *ev, int nev,
		int32_t *out, long frames, int buffer);

/* Bare DSP helpers, exported for unit tests */
unsigned a2o_p2i(int pitch);
int a2o_hermite(const int16_t *d, unsigned ph);
int a2o_lerp(const int16_t *d, unsigned ph);
int a2o_noise(uint32_t *state);
int a2o_f12_coeff(int cutoff_value_8_24, int samplerate);
void a2o_f12_coeff_array(const int *cutoff_values, int n, int samplerate, int *out);

#ifdef __cplusplus
}
#endif
#endif
