/*
 * a2cu_units.c - drop-in replacements for Audiality 2's hot-path voice units.
 *
 * Exports the A2_unitdesc symbols the reference's a2_core_units[] table
 * (src/audiality2.c:183-207) links against:
 *
 *     a2_wtosc_unitdesc  a2_panmix_unitdesc  a2_filter12_unitdesc
 *     a2_waveshaper_unitdesc  a2_fm1 .. a2_fm4, a2_fm3p, a2_fm4p, a2_fm2r,
 *     a2_fm4r _unitdesc,  a2_inline_unitdesc (bus bracketing) and
 *     a2_fbdelay_unitdesc, a2_limiter_unitdesc, a2_dcblock_unitdesc,
 *     a2_dc_unitdesc (the bus / song-level effects that follow the mix-down,
 *     on the host each costs a device round trip per call)
 *
 * The host (A2S compiler, VM, event scheduler, voice tree, drivers) is the
 * reference's own code, unmodified; it calls Initialize / write / Process /
 * Deinitialize exactly as it calls its built-in units (src/core.c:163-308,
 * 143-149, 1875-1876, 318-327). These callbacks do NO DSP: they record what
 * the host asked for into the CUDA engine (include/a2cu.h, block mode) and
 * make the result visible in host buffers only where an un-replaced host unit
 * is about to read it. There is no CPU fallback: a voice structure without a
 * kernel fails to instantiate (A2_NOTIMPLEMENTED), it is never rendered on
 * the host.
 *
 * This file compiles against the reference's own headers where they lie
 * (include/ for the public plug-in API, src/internals.h + src/units/inline.h
 * for the two private structs the reference's own inline unit also uses:
 * A2_voice and A2_inline) - see plugin/Makefile. Nothing is copied.
 *
 * Classification of a voice (at its first Process call, when the unit chain
 * is complete; chains never change afterwards, src/core.c:155-161):
 *   LEAF  every audio unit is ours, no inline, and the structure has a fused
 *         kernel: one device voice slot, the whole chain runs in one kernel
 *         (host units without audio I/O, e.g. `env`, only write controls)
 *   GEN   anything else: each of our units runs as its own device call on
 *         the voice's scratch channels, which live in a device bus row
 *         (`inline` = the bus its sub-voices mixed into); host units in the
 *         chain (xinsert, fbdelay, dcblock, ...) see materialised host
 *         buffers, and what they write is uploaded before our next unit.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "internals.h"
#include "inline.h"
#include "a2cu.h"
#include "a2cu_units.h"

#define A2CU_MAXCHAIN	12

typedef struct A2CU_ctx A2CU_ctx;
typedef struct A2CU_voice A2CU_voice;

/* Our instance data; the A2_inline layout must stay first for core.c:1763 */
typedef struct A2CU_unit
{
	A2_inline	il;		/* header (+ voice, state for inline) */
	A2CU_ctx	*cx;
	A2CU_voice	*voice;
	int		kind;		/* A2CU_* or 0 for inline */
	int		index;		/* position among the voice's units */
	unsigned	flags;		/* A2_PROCADD */
	unsigned	substart;
	int		init_transpose;
	int		*transpose;
	int		pm;		/* device panmix instance (GEN voices) */
	int		gu;		/* device generic unit (GEN voices) */
	int		bus;		/* inline: device bus of this fragment */
	unsigned	bus_serial;
	/*
	 * First unit of a LEAF voice: everything its Process() needs, so the
	 * per-segment path touches this block only (the host has just read
	 * u->Process from it) and not the voice record or the last unit.
	 */
	int		leaf;		/* 1: fused leaf voice, fields below valid */
	int		pool, slot;
	int32_t		**leaf_outputs;	/* where the voice's last unit is wired */
	int		leaf_nout;
	unsigned	proc_serial;	/* fragment of the last recorded segment */
	unsigned	cursor;		/* frame where the next segment starts */
	A2_voice	*hv;		/* the host voice this unit belongs to */
	/*
	 * History prefetcher: the first unit that was recorded right after this one in the
	 * previous fragment (the walk order is stable from fragment to fragment), and this
	 * voice's second unit. Only ever used as prefetch addresses.
	 */
	struct A2CU_unit *hint;
	struct A2CU_unit *follower;
} A2CU_unit;

typedef struct A2CU_pending
{
	int		unit, reg, value;
	unsigned	start, dur;
} A2CU_pending;

enum { VC_NEW = 0, VC_LEAF, VC_GEN, VC_BAD };

struct A2CU_voice
{
	A2_vmstate	*vms;
	int		refs;
	int		nunits;
	A2CU_unit	*units[A2CU_MAXCHAIN];
	int		cls;
	int		pool, slot;
	A2CU_pending	*pend;
	int		npend, cpend;
	/* GEN voices: where the scratch channels of the current segment live */
	int		scr_bus;	/* device bus row, or -1 */
	unsigned	scr_serial;	/* fragment scr_bus belongs to */
	int		scr_nch;
	int		on_device;
	uint64_t	run;		/* a2cu_block_run handle (bus commands) */
	/* Frame (fragment-relative) where the voice's next segment starts */
	unsigned	cursor;
	unsigned	cursor_serial;
};

typedef struct A2CU_wavemap { A2_wave *w; int id; const void *d0; unsigned s0; } A2CU_wavemap;
typedef struct A2CU_owner { int32_t **outputs; A2CU_unit *il; } A2CU_owner;
/* Host-only bus (no inline of ours owns it, e.g. the master) fed by voices */
typedef struct A2CU_orphan { int32_t **outputs; int bus, nch; } A2CU_orphan;
#define A2CU_MAXORPHANS	16

struct A2CU_ctx
{
	A2_config	*cfg;
	A2_state	*st;
	a2cu_engine	*eng;
	int		refs;
	unsigned	serial;		/* fragment counter */
	unsigned	last_fragstart;
	int		started;
	A2CU_wavemap	*waves;
	int		nwaves;
	A2CU_owner	owners[A2_NESTLIMIT];
	int		nowners;
	A2CU_orphan	orphans[A2CU_MAXORPHANS];
	int		norphans;
	A2CU_voice	*last_voice;	/* voice being populated */
	A2_voice	*look;		/* walk_prefetch: sibling whose lines are in flight */
	struct A2CU_unit *prev_first;	/* first unit recorded by the previous Process() */
	A2CU_ctx	*next;
};

static A2CU_ctx *contexts = NULL;
static int a2cu_device = 0;
static int a2cu_trace = -1;
#define TRACE(...)							\
	do {								\
		if(a2cu_trace < 0)					\
			a2cu_trace = getenv("A2CU_TRACE") ? 1 : 0;	\
		if(a2cu_trace)						\
			fprintf(stderr, "a2cu: " __VA_ARGS__);		\
	} while(0)

static int is_ours(const A2_unitdesc *d);

/*
 * A2CU_NULL=1: measurement aid, never a render mode. Every Process() / write
 * callback returns at once (inline still recurses into the host's walk), so a
 * run costs what the unmodified host spends on its own VM, event handling and
 * voice tree walk plus the bare indirect calls - the floor no unit library can
 * get under from behind the A2_unitdesc boundary (profiles/cfg5_k2trance.py
 * reports it next to the real run). The output is silence.
 */
static int a2cu_null = -1;
static inline int null_on(void)
{
	if(a2cu_null < 0)
		a2cu_null = getenv("A2CU_NULL") ? 1 : 0;
	return a2cu_null;
}

/*
 * A2CU_STATS=1: time spent inside the plug-in's callbacks (TSC ticks), printed
 * when the last engine state closes. Shows how much of a drop-in render is the
 * host's own VM + tree walk and how much is our recording.
 */
static int a2cu_stats = -1;
static unsigned long long st_proc_t, st_proc_n, st_write_t, st_write_n,
		st_inline_t, st_inline_n, st_t0;
static inline unsigned long long tsc(void)
{
#if defined(__x86_64__)
	unsigned lo, hi;
	__asm__ __volatile__("rdtsc" : "=a"(lo), "=d"(hi));
	return ((unsigned long long)hi << 32) | lo;
#else
	return 0;
#endif
}
static inline int stats_on(void)
{
	if(a2cu_stats < 0)
	{
		a2cu_stats = getenv("A2CU_STATS") ? 1 : 0;
		st_t0 = tsc();
	}
	return a2cu_stats;
}


/*---------------------------------------------------------
	Context (one per engine state), OpenState/CloseState
---------------------------------------------------------*/

static A2_errors a2cu_OpenState(A2_config *cfg, void **statedata)
{
	A2CU_ctx *cx;
	for(cx = contexts; cx; cx = cx->next)
		if(cx->cfg == cfg)
			break;
	if(!cx)
	{
		const char *dev = getenv("A2CU_DEVICE");
		if(dev)
			a2cu_device = atoi(dev);
		if(!(cx = (A2CU_ctx *)calloc(1, sizeof(A2CU_ctx))))
			return A2_OOMEMORY;
		cx->cfg = cfg;
		cx->st = ((A2_interface_i *)cfg->interface)->state;
		cx->eng = a2cu_open(a2cu_device, cfg->samplerate, 2);
		if(!cx->eng)
		{
			fprintf(stderr, "a2cu_units: %s\n", a2cu_last_error());
			free(cx);
			return A2_DEVICEOPEN;
		}
		a2cu_set_noise_state_ptr(cx->eng, &cx->st->noisestate);
		cx->next = contexts;
		contexts = cx;
	}
	++cx->refs;
	*statedata = cx;
	return A2_OK;
}

static void a2cu_CloseState(void *statedata)
{
	A2CU_ctx *cx = (A2CU_ctx *)statedata, **p;
	if(!cx || --cx->refs)
		return;
	for(p = &contexts; *p; p = &(*p)->next)
		if(*p == cx)
		{
			*p = cx->next;
			break;
		}
	if(stats_on())
	{
		unsigned long long total = tsc() - st_t0;
		fprintf(stderr, "a2cu plug-in stats (TSC ticks, %% of the time since the first callback): "
				"Process %llu calls %.1f%% (%.0f ticks/call), "
				"write %llu calls %.1f%% (%.0f ticks/call), "
				"inline (excl. recursion, incl. flush/sync) %llu calls %.1f%%\n",
				st_proc_n, 100.0 * st_proc_t / total,
				st_proc_n ? (double)st_proc_t / st_proc_n : 0.0,
				st_write_n, 100.0 * st_write_t / total,
				st_write_n ? (double)st_write_t / st_write_n : 0.0,
				st_inline_n, 100.0 * (double)(long long)st_inline_t / total);
	}
	a2cu_close(cx->eng);
	free(cx->waves);
	free(cx);
}

/* New fragment? (a2_AudioCallback advances now_fragstart, core.c:1964-1973) */
static inline void frag_check(A2CU_ctx *cx)
{
	if(cx->started && (cx->last_fragstart == cx->st->now_fragstart))
		return;
	cx->started = 1;
	cx->last_fragstart = cx->st->now_fragstart;
	++cx->serial;
	cx->nowners = 0;
	cx->norphans = 0;
	a2cu_block_begin(cx->eng);
}

static A2CU_unit *owner_find(A2CU_ctx *cx, int32_t **outputs)
{
	int i;
	for(i = cx->nowners - 1; i >= 0; --i)
		if(cx->owners[i].outputs == outputs)
			return cx->owners[i].il;
	return NULL;
}

/*
 * Device copy of a host wave. The cache is keyed on what the oscillator would
 * see through a2_GetWave at this moment - the A2_wave object, its level 0 data
 * pointer and length - so a wave that was released and re-created at the same
 * address, or re-rendered into new buffers, gets a fresh upload instead of
 * the stale copy, and a wave unloaded while in use (size[0] == 0,
 * waves.c:718-724, wtosc_check_unloaded wtosc.c:168-183) is unloaded on the
 * device too. Called on every `w` register write.
 */
static int wave_id(A2CU_ctx *cx, int handle)
{
	int i;
	A2_wave *w = a2_GetWave(cx->cfg->interface, handle);
	const void *d0 = NULL;
	unsigned s0 = 0;
	if(!w)
		return -1;
	if(w->type == A2_WWAVE || w->type == A2_WMIPWAVE)
	{
		d0 = w->d.wave.data[0];
		s0 = w->d.wave.size[0];
	}
	for(i = 0; i < cx->nwaves; ++i)
		if(cx->waves[i].w == w)
		{
			if(cx->waves[i].d0 == d0 && cx->waves[i].s0 == s0)
				return cx->waves[i].id;
			/* same object, other contents: retire the device copy */
			a2cu_wave_unload(cx->eng, cx->waves[i].id);
			cx->waves[i] = cx->waves[--cx->nwaves];
			break;
		}
	if((w->type == A2_WWAVE || w->type == A2_WMIPWAVE) && !s0)
		return -1;		/* unloaded: the oscillator turns off */
	{
		A2CU_wavemap *nw = (A2CU_wavemap *)realloc(cx->waves,
				sizeof(A2CU_wavemap) * (cx->nwaves + 1));
		int id;
		if(!nw)
			return -1;
		cx->waves = nw;
		id = a2cu_wave_upload_prepared(cx->eng, w->type, w->period,
				w->flags & A2_LOOPED,
				(const int16_t *const *)w->d.wave.data,
				w->d.wave.size);
		if(id < 0)
			return -1;
		cx->waves[cx->nwaves].w = w;
		cx->waves[cx->nwaves].id = id;
		cx->waves[cx->nwaves].d0 = d0;
		cx->waves[cx->nwaves].s0 = s0;
		++cx->nwaves;
		return id;
	}
}


/*---------------------------------------------------------
	Voice records
---------------------------------------------------------*/

static A2CU_voice *voice_for(A2CU_ctx *cx, A2_vmstate *vms, int first)
{
	A2CU_voice *v = cx->last_voice;
	if(!first && v && (v->vms == vms) && (v->cls == VC_NEW))
		return v;
	if(!(v = (A2CU_voice *)calloc(1, sizeof(A2CU_voice))))
		return NULL;
	v->vms = vms;
	v->pool = v->slot = -1;
	v->scr_bus = -1;
	cx->last_voice = v;
	return v;
}

static void voice_release(A2CU_ctx *cx, A2CU_voice *v)
{
	if(--v->refs)
		return;
	if(v->cls == VC_LEAF && v->slot >= 0)
		a2cu_pool_free(cx->eng, v->pool, v->slot);
	if(cx->last_voice == v)
		cx->last_voice = NULL;
	free(v->pend);
	free(v);
}

/* Bus commands recorded from here on belong to voice 'v' */
static inline void select_run(A2CU_ctx *cx, A2CU_voice *v)
{
	v->run = a2cu_block_run(cx->eng, a2_voice_from_vms(v->vms)->nestlevel,
			v->run);
}

static void emit_write(A2CU_ctx *cx, A2CU_voice *v, A2CU_unit *au, int reg,
		int value, unsigned frame, unsigned start, unsigned dur)
{
	if(v->cls == VC_LEAF)
	{
		if(au->kind == A2CU_WTOSC && reg == 0)
			value = wave_id(cx, value >> 16) << 16;
		a2cu_block_write(cx->eng, v->pool, v->slot, au->index, reg,
				value, *au->transpose, frame, start, dur);
	}
	else if(v->cls == VC_GEN && au->kind == A2CU_PANMIX && au->pm >= 0)
	{
		select_run(cx, v);
		a2cu_block_pm_write(cx->eng, au->pm, reg, value, start, dur);
	}
	else if(v->cls == VC_GEN && au->gu >= 0)
	{
		select_run(cx, v);
		if(au->kind == A2CU_WTOSC && reg == 0)
			value = wave_id(cx, value >> 16) << 16;
		a2cu_block_unit_write(cx->eng, au->gu, reg, value,
				*au->transpose, start, dur);
	}
}

static inline int has_audio_io(const A2_unit *u)
{
	return u->ninputs || u->noutputs;
}

/* Next unit of the chain that touches audio buffers (skips env & co) */
static A2_unit *next_audio_unit(A2_unit *u)
{
	for(u = u->next; u; u = u->next)
		if(is_ours(u->descriptor) || has_audio_io(u))
			return u;
	return NULL;
}

static void leaf_follower_process(A2_unit *u, unsigned offset, unsigned frames)
{
}

/* Decide what this voice is, now that its unit chain is complete. */
static void classify(A2CU_ctx *cx, A2CU_voice *v, unsigned frame)
{
	A2_voice *hv = a2_voice_from_vms(v->vms);
	A2_unit *u;
	a2cu_unitspec chain[A2CU_MAXCHAIN];
	int n = 0, fused = 1, i;
	{
		/*
		 * A fused (LEAF) voice records its whole chain at its first unit's
		 * Process(): that is only the reference's order of effects while
		 * no host unit sits BEHIND one of ours. A host unit with audio I/O
		 * needs the buffers; a control-only host unit (env) in mid-chain
		 * writes our registers between two of our Process() calls
		 * (core.c:1875-1876, env.c:120-137), which only the unit-by-unit
		 * (GEN) path reproduces.
		 */
		int seen_ours = 0;
		for(u = hv->units; u; u = u->next)
			if(is_ours(u->descriptor))
				seen_ours = 1;
			else if(has_audio_io(u) || seen_ours)
				fused = 0;
	}
	for(i = 0; i < v->nunits && n < A2CU_MAXCHAIN; ++i, ++n)
	{
		A2_unit *h = &v->units[i]->il.header;
		if(!v->units[i]->kind)
			fused = 0;		/* inline */
		chain[n].kind = v->units[i]->kind;
		chain[n].ninputs = h->ninputs;
		chain[n].noutputs = h->noutputs;
		chain[n].add = (v->units[i]->flags & A2_PROCADD) ? 1 : 0;
		chain[n].wireout = h->outputs == hv->outputs;
	}
	v->cls = VC_BAD;
	if(fused && n && a2cu_chain_supported(chain, n))
	{
		v->pool = a2cu_pool_open(cx->eng, chain, n);
		if(v->pool < 0)
		{
			a2r_Error(cx->st, A2_NOTIMPLEMENTED, a2cu_last_error());
			return;
		}
		v->slot = a2cu_pool_alloc(cx->eng, v->pool);
		if(v->slot < 0)
		{
			a2r_Error(cx->st, A2_OOMEMORY, "a2cu: voice slot");
			return;
		}
		v->cls = VC_LEAF;
		for(i = 0; i < v->nunits; ++i)
			a2cu_block_init(cx->eng, v->pool, v->slot, i,
					v->units[i]->init_transpose, frame,
					v->units[i]->substart);
		/* the first unit records the segment for the whole fused chain */
		for(i = 1; i < v->nunits; ++i)
			v->units[i]->il.header.Process = leaf_follower_process;
		v->units[0]->leaf = 1;
		v->units[0]->follower = v->nunits > 1 ? v->units[1] : NULL;
		v->units[0]->pool = v->pool;
		v->units[0]->slot = v->slot;
		v->units[0]->leaf_outputs =
				v->units[v->nunits - 1]->il.header.outputs;
		v->units[0]->leaf_nout =
				v->units[v->nunits - 1]->il.header.noutputs;
	}
	else
	{
		/* GEN voice: every DSP unit of ours gets its own device state */
		select_run(cx, v);
		for(i = 0; i < v->nunits; ++i)
		{
			A2CU_unit *au = v->units[i];
			A2_unit *h = &au->il.header;
			if(au->kind == 0)
				continue;
			if(au->kind == A2CU_PANMIX)
			{
				if((au->pm = a2cu_pm_alloc(cx->eng)) < 0)
				{
					a2r_Error(cx->st, A2_OOMEMORY,
							a2cu_last_error());
					return;
				}
				continue;
			}
			au->gu = a2cu_unit_alloc(cx->eng, au->kind, h->ninputs,
					h->noutputs);
			if(au->gu < 0)
			{
				a2r_Error(cx->st, A2_NOTIMPLEMENTED,
						a2cu_last_error());
				return;
			}
			a2cu_block_unit_init(cx->eng, au->gu,
					au->init_transpose, au->substart);
		}
		v->cls = VC_GEN;
	}
	/* Control writes that arrived before the first Process() call */
	for(i = 0; i < v->npend; ++i)
		emit_write(cx, v, v->units[v->pend[i].unit], v->pend[i].reg,
				v->pend[i].value, frame, v->pend[i].start,
				v->pend[i].dur);
	v->npend = 0;
}


/*---------------------------------------------------------
	Callbacks shared by all replaced units
---------------------------------------------------------*/

static A2_errors unit_init(A2_unit *u, A2_vmstate *vms, void *statedata,
		unsigned flags, int kind)
{
	A2CU_unit *au = (A2CU_unit *)u;
	A2CU_ctx *cx = (A2CU_ctx *)statedata;
	A2_voice *hv = a2_voice_from_vms(vms);
	A2CU_voice *v;
	int j, nregs = 0;
	if(!cx)
		return A2_INTERNAL + 900;
	/* First unit of a voice: v->units is still empty (core.c:299-303) */
	if(!(v = voice_for(cx, vms, hv->units == NULL)))
		return A2_OOMEMORY;
	if(v->nunits >= A2CU_MAXCHAIN)
		return A2_NOTIMPLEMENTED;
	au->cx = cx;
	au->voice = v;
	au->kind = kind;
	au->index = v->nunits;
	au->flags = flags;
	au->substart = vms->waketime & 0xff;
	au->transpose = vms->r + R_TRANSPOSE;
	au->init_transpose = *au->transpose;
	au->pm = -1;
	au->gu = -1;
	au->bus = -1;
	au->leaf = 0;
	au->proc_serial = 0;
	au->cursor = 0;
	au->hv = hv;
	au->hint = NULL;
	au->follower = NULL;
	au->bus_serial = 0;
	v->units[v->nunits++] = au;
	++v->refs;
	/* Register defaults, as each unit's Initialize() leaves them */
	if(u->descriptor->registers)
		while(u->descriptor->registers[nregs].name)
			++nregs;
	for(j = 0; j < nregs; ++j)
		u->registers[j] = 0;
	if(kind == A2CU_PANMIX)
		u->registers[0] = 65536;	/* panmix.c:264 */
	else if(kind == A2CU_FILTER12)
		u->registers[2] = 65536;	/* filter12.c:194 */
	else if(kind == A2CU_LIMITER)
	{
		u->registers[0] = 64 << 16;	/* limiter.c:165-166 */
		u->registers[1] = 1 << 16;
	}
	else if(kind == A2CU_DCBLOCK)
		u->registers[0] = -5 << 16;	/* dcblock.c:127 */
	else if(kind == A2CU_DC)
		u->registers[1] = 1 << 16;	/* dc.c:164: LINEAR */
	return A2_OK;
}

static void unit_deinit(A2_unit *u)
{
	A2CU_unit *au = (A2CU_unit *)u;
	if(au->cx && (au->cx->prev_first == au))
		au->cx->prev_first = NULL;	/* the block is about to be recycled */
	if(au->pm >= 0)
		a2cu_pm_free(au->cx->eng, au->pm);
	if(au->gu >= 0)
		a2cu_unit_free(au->cx->eng, au->gu);
	if(au->voice)
		voice_release(au->cx, au->voice);
}

static void unit_write_body(A2_unit *u, int reg, int value, unsigned start,
		unsigned dur);

static void unit_write(A2_unit *u, int reg, int value, unsigned start,
		unsigned dur)
{
	if(null_on())
		return;
	if(stats_on())
	{
		unsigned long long t = tsc();
		unit_write_body(u, reg, value, start, dur);
		st_write_t += tsc() - t;
		++st_write_n;
	}
	else
		unit_write_body(u, reg, value, start, dur);
}

static void unit_write_body(A2_unit *u, int reg, int value, unsigned start,
		unsigned dur)
{
	A2CU_unit *au = (A2CU_unit *)u;
	A2CU_voice *v = au->voice;
	A2CU_ctx *cx = au->cx;
	if(v->cls == VC_NEW)
	{
		if(v->npend == v->cpend)
		{
			int nc = v->cpend ? v->cpend * 2 : 16;
			A2CU_pending *np = (A2CU_pending *)realloc(v->pend,
					sizeof(A2CU_pending) * nc);
			if(!np)
				return;
			v->pend = np;
			v->cpend = nc;
		}
		v->pend[v->npend].unit = au->index;
		v->pend[v->npend].reg = reg;
		v->pend[v->npend].value = value;
		v->pend[v->npend].start = start;
		v->pend[v->npend].dur = dur;
		++v->npend;
		return;
	}
	frag_check(cx);
	/*
	 * Writes take effect where the voice's next Process() segment starts:
	 * the VM runs at "now" = the end of the previous segment
	 * (core.c:1847-1880), and so do host units that write controls from
	 * inside their own Process() (env.c:120-137) ahead of ours.
	 */
	{
		unsigned frame;
		if(v->cls == VC_LEAF)
		{
			A2CU_unit *u0 = v->units[0];
			frame = u0->proc_serial == cx->serial ? u0->cursor : 0;
		}
		else
			frame = v->cursor_serial == cx->serial ? v->cursor : 0;
		emit_write(cx, v, au, reg, value, frame, start, dur);
	}
}

#define	WRITE_CB(n)							\
static void unit_write##n(A2_unit *u, int v, unsigned s, unsigned d)	\
{									\
	unit_write(u, n, v, s, d);					\
}
WRITE_CB(0) WRITE_CB(1) WRITE_CB(2) WRITE_CB(3) WRITE_CB(4) WRITE_CB(5)
WRITE_CB(6) WRITE_CB(7) WRITE_CB(8) WRITE_CB(9) WRITE_CB(10) WRITE_CB(11)
WRITE_CB(12)

/* Make 'bus' visible in the host buffers 'bufs' (before a host unit runs) */
static void materialize(A2CU_ctx *cx, int bus, int nch, unsigned offset,
		unsigned frames, int32_t **bufs, int add)
{
	TRACE("materialize bus %d nch %d [%u,+%u) add %d\n", bus, nch, offset,
			frames, add);
	if(a2cu_block_download(cx->eng, bus, nch > 2 ? 2 : nch, offset, frames,
			bufs, add))
	{
		TRACE("download failed: %s\n", a2cu_last_error());
		a2r_Error(cx->st, A2_READ, a2cu_last_error());
	}
}

/* New segment of a GEN voice: nothing has written its scratch channels yet */
static inline void seg_begin(A2CU_ctx *cx, A2CU_voice *v, A2CU_unit *au)
{
	if(v->scr_serial != cx->serial)
	{
		v->scr_serial = cx->serial;
		v->scr_bus = -1;
		v->on_device = 0;
	}
	if(au->index == 0)
	{
		v->on_device = 0;
		v->scr_nch = 0;
	}
}

static inline void seg_mark(A2CU_ctx *cx, A2CU_voice *v, unsigned offset,
		unsigned frames)
{
	v->cursor = offset + frames;
	v->cursor_serial = cx->serial;
}

/* A host unit follows and the scratch channels are on the device: hand over */
static void handover(A2CU_ctx *cx, A2CU_voice *v, A2_unit *u, unsigned offset,
		unsigned frames)
{
	A2_unit *n;
	if(!v->on_device)
		return;
	n = next_audio_unit(u);
	if(!n || is_ours(n->descriptor))
		return;
	if(v->scr_nch)
		materialize(cx, v->scr_bus, v->scr_nch, offset, frames,
				n->inputs, 0);
	v->on_device = 0;
}

static void unit_process_body(A2_unit *u, unsigned offset, unsigned frames);

/* Process() of every replaced DSP unit (never inline) */
static void unit_process(A2_unit *u, unsigned offset, unsigned frames)
{
	if(null_on())
		return;
	if(stats_on())
	{
		unsigned long long t = tsc();
		unit_process_body(u, offset, frames);
		st_proc_t += tsc() - t;
		++st_proc_n;
	}
	else
		unit_process_body(u, offset, frames);
}

/*
 * The host's tree walk (a2_ProcessVoices, core.c:1883-1896) is a pointer chase
 * over voice structs (~1.5 KB each: `next` in the first line, `units` more
 * than a KB further) and unit blocks: with tens of thousands of voices every
 * one of those lines misses the CPU caches, for the host's own code and for
 * ours. While voice k is being recorded we therefore prefetch the unit block
 * of sibling k+1 (whose struct lines were requested one call earlier) and the
 * struct lines of sibling k+2. Prefetches never fault; a stale pointer only
 * wastes a line.
 */
static inline void prefetch_voice(const A2_voice *v)
{
	__builtin_prefetch(v);
	__builtin_prefetch(&v->s.waketime);
	__builtin_prefetch(&v->units);
}

/*
 * Unit blocks two recordings ahead, through our own hints (the sibling pointers above
 * reach the voice structs, but a voice's unit block address is only known once its
 * struct has arrived). 'h1' was requested one call ago, so its hint can be read now.
 */
static inline void hint_prefetch(A2CU_ctx *cx, A2CU_unit *au)
{
	A2CU_unit *h1 = au->hint, *h2;
	if(cx->prev_first)
		cx->prev_first->hint = au;
	cx->prev_first = au;
	if(!h1)
		return;
	if(h1->follower)
		__builtin_prefetch(h1->follower);
	if((h2 = h1->hint))
	{
		__builtin_prefetch(h2);
		__builtin_prefetch((const char *)h2 + 64);
		__builtin_prefetch((const char *)h2 + 128);
	}
}

static inline void walk_prefetch(A2CU_ctx *cx, const A2_voice *hv)
{
	A2_voice *n1 = hv->next;
	if(n1 && (cx->look == n1))
	{
		A2_voice *n2 = n1->next;
		const char *u1 = (const char *)n1->units;
		if(u1)
		{
			__builtin_prefetch(u1);
			__builtin_prefetch(u1 + 64);
			__builtin_prefetch(u1 + 128);
		}
		if(n2)
			prefetch_voice(n2);
		cx->look = n2;
	}
	else
	{
		if(n1)
			prefetch_voice(n1);
		cx->look = n1;
	}
}

static void leaf_process(A2CU_ctx *cx, A2CU_unit *au, unsigned offset,
		unsigned frames)
{
	A2CU_unit *owner;
	walk_prefetch(cx, au->hv);
	hint_prefetch(cx, au);
	au->proc_serial = cx->serial;
	au->cursor = offset + frames;
	owner = owner_find(cx, au->leaf_outputs);
	TRACE("leaf proc slot %d [%u,+%u) owner %p\n", au->slot, offset, frames,
			(void *)owner);
	if(owner)
		a2cu_block_proc(cx->eng, au->pool, au->slot, offset, frames,
				owner->bus);
	else
	{
		/*
		 * The voice mixes into a bus no inline of ours owns (a voice
		 * started before the root driver's INITV inherits the master
		 * bus, core.c:479-480): give that host bus a device shadow,
		 * added back when the outermost inline returns.
		 */
		int i, bus = -1;
		int32_t **outs = au->leaf_outputs;
		for(i = 0; i < cx->norphans; ++i)
			if(cx->orphans[i].outputs == outs)
				bus = cx->orphans[i].bus;
		if(bus < 0)
		{
			if(cx->norphans >= A2CU_MAXORPHANS)
			{
				a2r_Error(cx->st, A2_NOTIMPLEMENTED,
						"a2cu: too many host buses");
				return;
			}
			bus = a2cu_block_bus(cx->eng);
			cx->orphans[cx->norphans].outputs = outs;
			cx->orphans[cx->norphans].bus = bus;
			cx->orphans[cx->norphans].nch = 0;
			i = cx->norphans++;
		}
		else
			for(i = 0; cx->orphans[i].outputs != outs; ++i)
				;
		if(au->leaf_nout > cx->orphans[i].nch)
			cx->orphans[i].nch = au->leaf_nout;
		a2cu_block_proc(cx->eng, au->pool, au->slot, offset, frames,
				bus);
	}
}

static void unit_process_body(A2_unit *u, unsigned offset, unsigned frames)
{
	A2CU_unit *au = (A2CU_unit *)u;
	A2CU_voice *v;
	A2CU_ctx *cx = au->cx;
	frag_check(cx);
	if(au->leaf)
	{
		leaf_process(cx, au, offset, frames);
		return;
	}
	v = au->voice;
	if(v->cls == VC_NEW)
	{
		classify(cx, v, offset);
		TRACE("classified voice %p: cls %d pool %d slot %d units %d\n",
				(void *)v, v->cls, v->pool, v->slot, v->nunits);
	}
	if(v->cls == VC_LEAF)
	{
		if(!au->index)
			leaf_process(cx, v->units[0], offset, frames);
		return;
	}
	if(v->cls != VC_GEN || (au->pm < 0 && au->gu < 0))
		return;
	/* One unit on the voice's device scratch channels */
	seg_begin(cx, v, au);
	seg_mark(cx, v, offset, frames);
	TRACE("gen unit kind %d pm %d gu %d [%u,+%u) on_device %d scr %d\n",
			au->kind, au->pm, au->gu, offset, frames, v->on_device,
			v->scr_bus);
	{
		int wire = u->outputs != u->inputs;
		int add = (au->flags & A2_PROCADD) ? 1 : 0;
		int out_bus = -1, host_out = 0;
		A2CU_unit *owner;
		if(!v->on_device)
		{
			/* Input (if any) was written by a host unit: bring it over */
			int nin = u->ninputs;
			if(add && !wire && u->noutputs > nin)
				nin = u->noutputs;
			if(v->scr_bus < 0)
				v->scr_bus = a2cu_block_bus(cx->eng);
			if(nin)
				a2cu_block_upload(cx->eng, v->scr_bus,
						nin > 2 ? 2 : nin, offset, frames,
						(const int32_t *const *)u->inputs);
			if(nin > v->scr_nch)
				v->scr_nch = nin;
			v->on_device = 1;
		}
		if(wire)
		{
			if((owner = owner_find(cx, u->outputs)))
				out_bus = owner->bus;
			else
			{
				/* Output is a host-only bus (e.g. the master) */
				out_bus = a2cu_block_bus(cx->eng);
				host_out = 1;
			}
		}
		/* after the upload above: a flush ends every command run */
		select_run(cx, v);
		if(au->pm >= 0)
			a2cu_block_pm_proc(cx->eng, au->pm, u->ninputs,
					u->noutputs, host_out ? 0 : add,
					v->scr_bus, wire ? out_bus : v->scr_bus,
					offset, frames);
		else
			a2cu_block_unit_proc(cx->eng, au->gu, add, wire,
					v->scr_bus, out_bus, offset, frames);
		if(host_out)
			materialize(cx, out_bus, u->noutputs, offset, frames,
					u->outputs, 1);
		if(!wire && u->noutputs > v->scr_nch)
			v->scr_nch = u->noutputs > 2 ? 2 : u->noutputs;
		handover(cx, v, u, offset, frames);
	}
}


/*---------------------------------------------------------
	inline (src/units/inline.c, src/core.c:1763-1776)
---------------------------------------------------------*/

static int any_nonzero(int32_t **bufs, int nch, unsigned offset, unsigned frames)
{
	int c;
	unsigned i;
	for(c = 0; c < nch; ++c)
		for(i = 0; i < frames; ++i)
			if(bufs[c][offset + i])
				return 1;
	return 0;
}

static void inline_process(A2_unit *u, unsigned offset, unsigned frames,
		int add)
{
	A2CU_unit *au = (A2CU_unit *)u;
	A2CU_voice *v = au->voice;
	A2CU_ctx *cx = au->cx;
	A2_unit *n;
	int i, dev_pre, prev_bus;
	int nch = u->noutputs > 2 ? 2 : u->noutputs;
	frag_check(cx);
	if(v->cls == VC_NEW)
		classify(cx, v, offset);
	seg_begin(cx, v, au);
	seg_mark(cx, v, offset, frames);
	if(au->bus_serial != cx->serial)
	{
		au->bus = a2cu_block_bus(cx->eng);
		au->bus_serial = cx->serial;
	}
	TRACE("inline %p [%u,+%u) add %d bus %d cls %d\n", (void *)au, offset,
			frames, add, au->bus, v->cls);
	/*
	 * Adding inline after units of ours: what they wrote is on the device;
	 * the host buffer then only collects what host units add below.
	 */
	dev_pre = add && v->on_device && (v->scr_bus >= 0) &&
			(v->scr_bus != au->bus);
	prev_bus = v->scr_bus;
	if(!add || dev_pre)
		for(i = 0; i < u->noutputs; ++i)
			memset(u->outputs[i] + offset, 0, frames * sizeof(int));
	if(au->il.voice->sub)
		prefetch_voice(au->il.voice->sub);
	/* Sub-voices that mix into u->outputs now mix into our device bus */
	if(cx->nowners < A2_NESTLIMIT)
	{
		cx->owners[cx->nowners].outputs = u->outputs;
		cx->owners[cx->nowners].il = au;
		++cx->nowners;
	}
	if(stats_on())
	{
		unsigned long long t = tsc();
		a2_inline_ProcessAdd(u, offset, frames);
		st_inline_t -= tsc() - t;	/* exclusive of the recursion */
	}
	else
		a2_inline_ProcessAdd(u, offset, frames);	/* the host's recursion */
	if(cx->nowners)
		--cx->nowners;
	frag_check(cx);
	if(!cx->nowners)
		for(i = 0; i < cx->norphans; ++i)	/* outermost inline */
			materialize(cx, cx->orphans[i].bus, cx->orphans[i].nch,
					offset, frames, cx->orphans[i].outputs,
					1);
	select_run(cx, v);	/* the recursion selected other voices */
	if(dev_pre)
		a2cu_block_bus_add(cx->eng, prev_bus, au->bus, offset, frames);
	n = next_audio_unit(u);
	if(v->cls == VC_GEN && n && is_ours(n->descriptor) &&
			(u->outputs == n->inputs))
	{
		/*
		 * Stays on the device. Host units of sub-voices wired to this
		 * bus (e.g. a song's `fbdelay * >`) added into the host
		 * buffer: bring that over.
		 */
		if(any_nonzero(u->outputs, nch, offset, frames))
		{
			int tmp = a2cu_block_bus(cx->eng);
			a2cu_block_upload(cx->eng, tmp, nch, offset, frames,
					(const int32_t *const *)u->outputs);
			select_run(cx, v);	/* the upload flushed */
			a2cu_block_bus_add(cx->eng, tmp, au->bus, offset,
					frames);
		}
		v->scr_bus = au->bus;
		v->scr_nch = nch;
		v->on_device = 1;
	}
	else
	{
		/* A host unit (or nothing of ours) follows: hand the sum over */
		materialize(cx, au->bus, u->noutputs, offset, frames,
				u->outputs, 1);
		v->on_device = 0;
	}
}

static void inline_timed(A2_unit *u, unsigned offset, unsigned frames, int add)
{
	if(null_on())
	{
		a2_inline_ProcessAdd(u, offset, frames);	/* the host's recursion only */
		return;
	}
	if(stats_on())
	{
		unsigned long long t = tsc();
		inline_process(u, offset, frames, add);
		st_inline_t += tsc() - t;
		++st_inline_n;
	}
	else
		inline_process(u, offset, frames, add);
}

static void inline_Process(A2_unit *u, unsigned offset, unsigned frames)
{
	inline_timed(u, offset, frames, 0);
}

static void inline_ProcessAdd(A2_unit *u, unsigned offset, unsigned frames)
{
	inline_timed(u, offset, frames, 1);
}

static A2_errors inline_Initialize(A2_unit *u, A2_vmstate *vms,
		void *statedata, unsigned flags)
{
	A2CU_unit *au = (A2CU_unit *)u;
	A2_errors res = unit_init(u, vms, statedata, flags, 0);
	if(res)
		return res;
	/* units/inline.c:26-39 */
	au->il.state = au->cx->st;
	au->il.voice = a2_voice_from_vms(vms);
	au->il.voice->noutputs = u->noutputs;
	au->il.voice->outputs = u->outputs;
	u->Process = (flags & A2_PROCADD) ? inline_ProcessAdd : inline_Process;
	return A2_OK;
}


/*---------------------------------------------------------
	Unit descriptors
---------------------------------------------------------*/

#define	INIT_CB(name, kind)						\
static A2_errors name##_Initialize(A2_unit *u, A2_vmstate *vms,	\
		void *statedata, unsigned flags)			\
{									\
	A2_errors res = unit_init(u, vms, statedata, flags, kind);	\
	u->Process = unit_process;					\
	return res;							\
}

INIT_CB(wtosc, A2CU_WTOSC)
INIT_CB(panmix, A2CU_PANMIX)
INIT_CB(filter12, A2CU_FILTER12)
INIT_CB(waveshaper, A2CU_WAVESHAPER)
INIT_CB(fm1, A2CU_FM1)
INIT_CB(fm2, A2CU_FM2)
INIT_CB(fm3, A2CU_FM3)
INIT_CB(fm4, A2CU_FM4)
INIT_CB(fm3p, A2CU_FM3P)
INIT_CB(fm4p, A2CU_FM4P)
INIT_CB(fm2r, A2CU_FM2R)
INIT_CB(fm4r, A2CU_FM4R)

INIT_CB(limiter, A2CU_LIMITER)
INIT_CB(dcblock, A2CU_DCBLOCK)
INIT_CB(dc, A2CU_DC)

/* units/fbdelay.c:176-225: default register values */
static A2_errors fbdelay_Initialize(A2_unit *u, A2_vmstate *vms,
		void *statedata, unsigned flags)
{
	A2_errors res = unit_init(u, vms, statedata, flags, A2CU_FBDELAY);
	u->Process = unit_process;
	if(res)
		return res;
	u->registers[0] = 400 << 16;
	u->registers[1] = 280 << 16;
	u->registers[2] = 320 << 16;
	u->registers[3] = 65536;
	u->registers[4] = 16384;
	u->registers[5] = 32768;
	u->registers[6] = 32768;
	return A2_OK;
}

/* Register names and order: the reference's A2_crdesc tables */
static const A2_crdesc wtosc_regs[] = {		/* wtosc.c:507-514 */
	{ "w", unit_write0 }, { "p", unit_write1 }, { "a", unit_write2 },
	{ "phase", unit_write3 }, { NULL, NULL }
};
static const A2_crdesc panmix_regs[] = {	/* panmix.c:298-303 */
	{ "vol", unit_write0 }, { "pan", unit_write1 }, { NULL, NULL }
};
static const A2_constdesc panmix_constants[] = {	/* panmix.c:305-311 */
	{ "CENTER", 0 }, { "LEFT", (-1) << 16 }, { "RIGHT", 1 << 16 },
	{ NULL, 0 }
};
static const A2_crdesc filter12_regs[] = {	/* filter12.c:231-239 */
	{ "cutoff", unit_write0 }, { "q", unit_write1 }, { "lp", unit_write2 },
	{ "bp", unit_write3 }, { "hp", unit_write4 }, { NULL, NULL }
};
static const A2_crdesc fbdelay_regs[] = {	/* fbdelay.c:286-296 */
	{ "fbdelay", unit_write0 }, { "ldelay", unit_write1 },
	{ "rdelay", unit_write2 }, { "drygain", unit_write3 },
	{ "fbgain", unit_write4 }, { "lgain", unit_write5 },
	{ "rgain", unit_write6 }, { NULL, NULL }
};
static const A2_crdesc limiter_regs[] = {	/* limiter.c:209-214 */
	{ "release", unit_write0 }, { "threshold", unit_write1 }, { NULL, NULL }
};
static const A2_crdesc dcblock_regs[] = {	/* dcblock.c:157-161 */
	{ "cutoff", unit_write0 }, { NULL, NULL }
};
static const A2_crdesc dc_regs[] = {		/* dc.c:249-254 */
	{ "value", unit_write0 }, { "mode", unit_write1 }, { NULL, NULL }
};
static const A2_constdesc dc_constants[] = {	/* dc.c:256-265 */
	{ "STEP", 0 << 16 }, { "LINEAR", 1 << 16 }, { NULL, 0 }
};
static const A2_crdesc waveshaper_regs[] = {	/* waveshaper.c:166-170 */
	{ "amount", unit_write0 }, { NULL, NULL }
};
/* fm.c:519-530 etc: phase, then (p, a, fb) per operator */
static const A2_crdesc fm_regs[] = {
	{ "phase", unit_write0 },
	{ "p", unit_write1 }, { "a", unit_write2 }, { "fb", unit_write3 },
	{ "p1", unit_write4 }, { "a1", unit_write5 }, { "fb1", unit_write6 },
	{ "p2", unit_write7 }, { "a2", unit_write8 }, { "fb2", unit_write9 },
	{ "p3", unit_write10 }, { "a3", unit_write11 }, { "fb3", unit_write12 },
	{ NULL, NULL }
};
static const A2_crdesc fm1_regs[] = {
	{ "phase", unit_write0 },
	{ "p", unit_write1 }, { "a", unit_write2 }, { "fb", unit_write3 },
	{ NULL, NULL }
};
static const A2_crdesc fm2_regs[] = {
	{ "phase", unit_write0 },
	{ "p", unit_write1 }, { "a", unit_write2 }, { "fb", unit_write3 },
	{ "p1", unit_write4 }, { "a1", unit_write5 }, { "fb1", unit_write6 },
	{ NULL, NULL }
};
static const A2_crdesc fm3_regs[] = {
	{ "phase", unit_write0 },
	{ "p", unit_write1 }, { "a", unit_write2 }, { "fb", unit_write3 },
	{ "p1", unit_write4 }, { "a1", unit_write5 }, { "fb1", unit_write6 },
	{ "p2", unit_write7 }, { "a2", unit_write8 }, { "fb2", unit_write9 },
	{ NULL, NULL }
};

#define	UNITDESC(sym, uname, uflags, regs, consts, mini, maxi, mino, maxo, init) \
const A2_unitdesc sym = {						\
	uname, uflags, regs, NULL, consts,				\
	mini, maxi, mino, maxo,						\
	sizeof(A2CU_unit), init, unit_deinit,				\
	a2cu_OpenState, a2cu_CloseState					\
};

/* I/O limits and flags as in the reference descriptors */
UNITDESC(a2_wtosc_unitdesc, "wtosc", 0, wtosc_regs, NULL, 0, 0, 1, 1,
		wtosc_Initialize)			/* wtosc.c:516-536 */
UNITDESC(a2_panmix_unitdesc, "panmix", 0, panmix_regs, panmix_constants,
		1, 2, 1, 2, panmix_Initialize)		/* panmix.c:313-333 */
UNITDESC(a2_filter12_unitdesc, "filter12", A2_MATCHIO, filter12_regs, NULL,
		1, 2, 1, 2, filter12_Initialize)	/* filter12.c:241-261 */
UNITDESC(a2_waveshaper_unitdesc, "waveshaper", A2_MATCHIO, waveshaper_regs,
		NULL, 1, 2, 1, 2, waveshaper_Initialize) /* waveshaper.c:173-193 */
UNITDESC(a2_fm1_unitdesc, "fm1", 0, fm1_regs, NULL, 0, 0, 1, 1, fm1_Initialize)
UNITDESC(a2_fm2_unitdesc, "fm2", 0, fm2_regs, NULL, 0, 0, 1, 1, fm2_Initialize)
UNITDESC(a2_fm3_unitdesc, "fm3", 0, fm3_regs, NULL, 0, 0, 1, 1, fm3_Initialize)
UNITDESC(a2_fm4_unitdesc, "fm4", 0, fm_regs, NULL, 0, 0, 1, 1, fm4_Initialize)
UNITDESC(a2_fm3p_unitdesc, "fm3p", 0, fm3_regs, NULL, 0, 0, 1, 1,
		fm3p_Initialize)
UNITDESC(a2_fm4p_unitdesc, "fm4p", 0, fm_regs, NULL, 0, 0, 1, 1,
		fm4p_Initialize)
UNITDESC(a2_fm2r_unitdesc, "fm2r", 0, fm2_regs, NULL, 0, 0, 1, 1,
		fm2r_Initialize)
UNITDESC(a2_fm4r_unitdesc, "fm4r", 0, fm_regs, NULL, 0, 0, 1, 1,
		fm4r_Initialize)
/* fbdelay.c:298-319 */
UNITDESC(a2_fbdelay_unitdesc, "fbdelay", 0, fbdelay_regs, NULL, 1, 2, 1, 2,
		fbdelay_Initialize)
/* limiter.c:216-238, dcblock.c:163-185, dc.c:267-289 */
UNITDESC(a2_limiter_unitdesc, "limiter", A2_MATCHIO, limiter_regs, NULL, 1, 2, 1, 2,
		limiter_Initialize)
UNITDESC(a2_dcblock_unitdesc, "dcblock", A2_MATCHIO, dcblock_regs, NULL, 1, 2, 1, 2,
		dcblock_Initialize)
UNITDESC(a2_dc_unitdesc, "dc", 0, dc_regs, dc_constants, 0, 0, 1, 2, dc_Initialize)
/* units/inline.c:49-69 */
UNITDESC(a2_inline_unitdesc, "inline", 0, NULL, NULL, 0, 0, 1, A2_MAXCHANNELS,
		inline_Initialize)

static int is_ours(const A2_unitdesc *d)
{
	return d->OpenState == a2cu_OpenState;
}


/*---------------------------------------------------------
	"cuda" audio driver (mirrors drivers/bufferdrv.c:28-110)
---------------------------------------------------------*/

static A2_errors cudad_Run(A2_audiodriver *driver, unsigned frames)
{
	A2_config *cfg = driver->driver.config;
	if(driver->Process)
		driver->Process(driver, frames);	/* a2_AudioCallback */
	else
	{
		int c;
		for(c = 0; c < cfg->channels; ++c)
			memset(driver->buffers[c], 0, sizeof(int32_t) * frames);
	}
	return A2_OK;
}

static void cudad_Lock(A2_audiodriver *driver) { }
static void cudad_Unlock(A2_audiodriver *driver) { }

static void cudad_Close(A2_driver *driver)
{
	A2_audiodriver *ad = (A2_audiodriver *)driver;
	A2_config *cfg = driver->config;
	if(ad->buffers)
	{
		int c;
		for(c = 0; c < cfg->channels; ++c)
			free(ad->buffers[c]);
		free(ad->buffers);
		ad->buffers = NULL;
	}
	ad->Run = NULL;
	ad->Lock = NULL;
	ad->Unlock = NULL;
}

static A2_errors cudad_Open(A2_driver *driver)
{
	A2_audiodriver *ad = (A2_audiodriver *)driver;
	A2_config *cfg = driver->config;
	int c;
	ad->Run = cudad_Run;
	ad->Lock = cudad_Lock;
	ad->Unlock = cudad_Unlock;
	if(!(ad->buffers = (int32_t **)calloc(cfg->channels, sizeof(int32_t *))))
		return A2_OOMEMORY;
	for(c = 0; c < cfg->channels; ++c)
		if(!(ad->buffers[c] = (int32_t *)calloc(cfg->buffer,
				sizeof(int32_t))))
		{
			cudad_Close(driver);
			return A2_OOMEMORY;
		}
	return A2_OK;
}

static A2_driver *cudad_new(A2_drivertypes type, const char *nameopts)
{
	A2_audiodriver *d = (A2_audiodriver *)calloc(1, sizeof(A2_audiodriver));
	if(!d)
		return NULL;
	d->driver.type = A2_AUDIODRIVER;
	d->driver.name = "cuda";
	d->driver.Open = cudad_Open;
	d->driver.Close = cudad_Close;
	return &d->driver;
}

int a2cu_RegisterDriver(void)
{
	return a2_RegisterDriver(A2_AUDIODRIVER, "cuda", cudad_new);
}
