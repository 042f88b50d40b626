/*
 * a2_oracle.c - CPU restatement of Audiality 2's per-voice DSP render path.
 *
 * TEST INFRASTRUCTURE ONLY - see a2_oracle.h.  Build: oracle/Makefile (port).
 * Compile with -fwrapv: the reference relies on two's complement wrap-around
 * of int (gcc x86-64 behaviour), e.g. a2_RunRamper (a2_dsp.h:152-155).
 *
 * Every function cites the reference file:line it restates.  The structure is
 * deliberately different from the reference (one file, plain state structs,
 * an explicit event list instead of the VM), the arithmetic is not: integer
 * expressions are kept operation for operation because parity is bit-exact.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "a2_oracle.h"

#define MIDDLEC		261.626f	/* include/a2_pitch.h:42 */
#define MAXPHINC	512		/* include/a2_waves.h:58 */
#define WAVEPERIOD	2048		/* include/a2_waves.h:71 */
#define WT_MAXLEN	(0x01000000 - A2O_WAVEPRE - A2O_WAVEPOST) /* wtosc.c:55 */

/* ------------------------------------------------------------------ */
/* a2_dsp.h                                                            */
/* ------------------------------------------------------------------ */

typedef struct ramp { int value, target, delta, timer; } ramp;

/* a2_dsp.h:37-42 */
int a2o_noise(uint32_t *state)
{
	uint32_t x = *state * 1566083941u + 1u;
	*state = x;
	return (int)(x * (x >> 16) >> 16);
}

/* a2_dsp.h:50-55 */
int a2o_lerp(const int16_t *d, unsigned ph)
{
	int i = ph >> 8, x = ph & 0xff;
	return (d[i] * (256 - x) + d[i + 1] * x) >> 8;
}

/* a2_dsp.h:64-74 */
int a2o_hermite(const int16_t *d, unsigned ph)
{
	int i = ph >> 8;
	int x = (ph & 0xff) << 7;
	int dm = d[i - 1], d0 = d[i], d1 = d[i + 1], d2 = d[i + 2];
	int c = (d1 - dm) >> 1;
	int a = (3 * (d0 - d1) + d2 - dm) >> 1;
	int b = dm - d0 + c - a;
	a = a * x >> 15;
	a = (a + b) * x >> 15;
	return d0 + ((a + c) * x >> 15);
}

/* a2_dsp.h:121-125 */
static void ramp_init(ramp *r, int v)
{
	r->value = r->target = v << 8;
	r->delta = r->timer = 0;
}

/* a2_dsp.h:128-149 */
static void ramp_prepare(ramp *r, int frames)
{
	if(!r->timer)
	{
		r->value = r->target;
		r->delta = 0;
	}
	else if(frames <= (r->timer >> 8))
	{
		r->delta = (int)(((int64_t)(r->target - r->value) << 8) /
				r->timer);
		r->timer -= frames << 8;
	}
	else
	{
		r->delta = (r->target - r->value) / frames;
		r->timer = 0;
	}
}

/* a2_dsp.h:152-155 */
static void ramp_run(ramp *r, int frames)
{
	r->value += r->delta * frames;
}

/* a2_dsp.h:161-170 */
static void ramp_set(ramp *r, int target, int start, int duration)
{
	r->target = target << 8;
	r->timer = duration + start;
	if(r->timer < 256)
		r->value = r->target;
	else
		r->value += r->delta * start >> 8;
}

/* ------------------------------------------------------------------ */
/* pitch.c                                                             */
/* ------------------------------------------------------------------ */

static unsigned ptab_base[64], ptab_coeff[64];
static int ptab_ready = 0;

/* pitch.c:70-96 */
static void ptab_build(void)
{
	unsigned i, b = 0x80000000u;
	for(i = 0; i < 64; ++i)
	{
		unsigned b2 = (double)0x80000000u *
				powf(2.0f, (i + 1) * (1.0f / 64)) + 0.5f;
		ptab_base[i] = b;
		ptab_coeff[i] = (b2 - b + 128) >> 8;
		b = b2;
	}
	ptab_ready = 1;
}

/*
 * pitch.c:57-67.  The final shift count (7 - oct) is outside 0..31 for
 * extreme pitches, which is undefined in C; the reference as built by gcc for
 * x86-64 gets the hardware's count & 31, so that is what we state.
 */
unsigned a2o_p2i(int pitch)
{
	int n = pitch & 0xffff;
	int oct = pitch >> 16;
	unsigned dph;
	if(!ptab_ready)
		ptab_build();
	dph = ptab_coeff[n >> 10] * (unsigned)(n & 0x3ff);
	dph >>= 2;
	dph += ptab_base[n >> 10];
	return dph >> ((7 - oct) & 31);
}

/* ------------------------------------------------------------------ */
/* waves.c (read side + preparation)                                   */
/* ------------------------------------------------------------------ */

typedef struct wave
{
	int		type;
	unsigned	flags;
	unsigned	period;
	int16_t		*data[A2O_MIPLEVELS];	/* incl. pads */
	unsigned	size[A2O_MIPLEVELS];	/* excl. pads */
	char		name[16];
} wave;

/* waves.c:90-106 */
static void wave_fix_pad(wave *w, int level)
{
	int16_t *d = w->data[level];
	unsigned size = w->size[level];
	if((w->flags & A2O_LOOPED) && size)
	{
		int i;
		memcpy(d, d + size, A2O_WAVEPRE * 2);
		for(i = 0; i < A2O_WAVEPOST; ++i)
			d[A2O_WAVEPRE + size + i] = d[A2O_WAVEPRE + i % size];
	}
	else
	{
		memset(d, 0, A2O_WAVEPRE * 2);
		memset(d + A2O_WAVEPRE + size, 0, A2O_WAVEPOST * 2);
	}
}

/* waves.c:59-87 (sizes), :108-132 (mip rendering) */
static int wave_prepare(wave *w, const int16_t *src, unsigned length)
{
	int i, levels = w->type == A2O_WMIPWAVE ? A2O_MIPLEVELS : 1;
	for(i = 0; i < levels; ++i)
	{
		unsigned size = (length + (1u << i) - 1) >> i;
		w->size[i] = size;
		w->data[i] = (int16_t *)calloc(A2O_WAVEPRE + size +
				A2O_WAVEPOST, sizeof(int16_t));
		if(!w->data[i])
			return -1;
	}
	memcpy(w->data[0] + A2O_WAVEPRE, src, length * sizeof(int16_t));
	wave_fix_pad(w, 0);
	for(i = 1; i < levels; ++i)
	{
		int s;
		int16_t *sd = w->data[i - 1] + A2O_WAVEPRE;
		int16_t *d = w->data[i] + A2O_WAVEPRE;
		for(s = 0; s < (int)w->size[i]; ++s)
			d[s] = (((int)sd[s * 2] << 1) + sd[s * 2 - 1] +
					sd[s * 2 + 1]) >> 2;
		wave_fix_pad(w, i);
	}
	return 0;
}

/* ------------------------------------------------------------------ */
/* Unit state                                                          */
/* ------------------------------------------------------------------ */

enum { OSC_OFF = 0, OSC_NOISE, OSC_NOMIP, OSC_MIP };

typedef struct st_wtosc	/* wtosc.c:66-80 */
{
	unsigned	dphase;
	uint64_t	phase;
	int		noise;
	int		p_ramping;
	ramp		p, a;
	int		wave;		/* index or -1 */
	int		mode;
} st_wtosc;

typedef struct st_panmix { ramp vol, pan; } st_panmix;	/* panmix.c:35-40 */

typedef struct st_f12		/* filter12.c:37-56 */
{
	ramp	cutoff, q;
	int	lp, bp, hp;
	int	f1;
	int	d1[2], d2[2];
} st_f12;

typedef struct st_fmop		/* fm.c:81-90 */
{
	ramp		a, fb, p;
	int		last_pitch;
	unsigned	phase, dphase;
	int		last;
} st_fmop;

typedef struct st_fm { int nops, osbits, par; st_fmop op[4]; } st_fm;

typedef struct st_ws { ramp amount; } st_ws;	/* waveshaper.c:42-46 */

typedef struct unit
{
	a2o_unitspec	spec;
	union {
		st_wtosc	osc;
		st_panmix	pm;
		st_f12		f12;
		st_fm		fm;
		st_ws		ws;
	} s;
} unit;

typedef struct voice
{
	int		alive;
	int		nunits;
	int		transpose;
	int		group;		/* -1: root bus */
	unit		u[A2O_MAXUNITS];
	/* event cursor for the current a2o_render() call */
	int		*evi;
	int		nev, evpos;
} voice;

/* fbdelay (units/fbdelay.c): the effect every song chains after its mix-down */
#define A2O_FBD_BUFSIZE	131072		/* A2FBD_BUFSIZE, fbdelay.c:26 */
typedef struct st_fbdelay
{
	int		fbdelay, ldelay, rdelay;	/* frames */
	int		drygain, fbgain, lgain, rgain;	/* 16:16 */
	int32_t		*lbuf, *rbuf;
	int		bufpos;
} st_fbdelay;

static int fbd_init(st_fbdelay *f, int samplerate);
static void fbd_write(st_fbdelay *f, int reg, int v, int samplerate);

typedef struct group
{
	st_panmix	pm;
	int		has_fbd;	/* { inline 0 *; fbdelay * *; panmix * > } */
	st_fbdelay	fbd;
	int		*evi;
	int		nev, evpos;
} group;

struct a2o_engine
{
	int		samplerate, channels;
	int		basepitch;
	uint32_t	msdur;
	uint32_t	noisestate;
	uint32_t	now_fragstart;		/* 24:8 */
	wave		*waves;
	int		nwaves;
	voice		*voices;
	int		nvoices, cvoices;
	group		*groups;
	int		ngroups;
	/* creation-order list of root children: >=0 voice, <0 ~group */
	int		*order;
	int		norder, corder;
	st_panmix	rootpm;
	int		*rootevi;
	int		rootnev, rootevpos;
	int16_t		fmsine[2049];
	const a2o_event	*ev;
	/* buses: scratch per nest level, [channel][frame] */
	int32_t		master[2][A2O_MAXFRAG];
	int32_t		rootbus[2][A2O_MAXFRAG];
	int32_t		groupbus[2][A2O_MAXFRAG];
	int32_t		vscratch[2][A2O_MAXFRAG];
};

/* ------------------------------------------------------------------ */
/* wtosc.c                                                             */
/* ------------------------------------------------------------------ */

/* wtosc.c:30-33 (A2_HIFI is hard-wired on, config.h:108) */
static int osc_inter(const int16_t *d, unsigned ph, unsigned dph)
{
	return a2o_hermite(d, ph) + a2o_hermite(d, ph + (dph >> 1));
}

/* wtosc.c:89-105 */
static void osc_run_pitch(st_wtosc *o, unsigned frames)
{
	unsigned lastv;
	ramp_prepare(&o->p, frames);
	if(o->dphase && (!o->p.timer && !o->p_ramping))
		return;
	lastv = o->p.value;
	ramp_run(&o->p, frames);
	o->p_ramping = o->p.delta;
	o->dphase = a2o_p2i((int)((lastv + (unsigned)o->p.value) >> 9));
}

/* wtosc.c:378-387 */
static void osc_set_phase(a2o_engine *e, st_wtosc *o, int ph, unsigned sst)
{
	if(o->wave < 0)
	{
		o->phase = 0;
		return;
	}
	ph += sst * (o->dphase >> 8) >> 8;
	o->phase = (uint64_t)((int64_t)ph * e->waves[o->wave].period << 8);
}

/* wtosc.c:390-423 */
static void osc_init(a2o_engine *e, voice *v, st_wtosc *o, unsigned substart)
{
	o->noise = 0;
	o->wave = -1;
	ramp_init(&o->a, 0);
	ramp_init(&o->p, v->transpose + e->basepitch);
	o->dphase = a2o_p2i(o->p.value >> 8);
	o->p_ramping = 0;
	osc_set_phase(e, o, 0, substart);
	o->mode = OSC_OFF;
}

/* wtosc.c:433-504 */
static void osc_write(a2o_engine *e, voice *v, st_wtosc *o, int reg, int val,
		unsigned start, unsigned dur)
{
	switch(reg)
	{
	  case A2O_W_WAVE:
	  {
		int wt = A2O_WOFF;
		int h = val >> 16;
		o->wave = (h >= 0 && h < e->nwaves) ? h : -1;
		if(o->wave >= 0)
			wt = e->waves[o->wave].type;
		if((wt == A2O_WWAVE || wt == A2O_WMIPWAVE) &&
				e->waves[o->wave].size[0] > WT_MAXLEN)
			wt = A2O_WOFF;
		switch(wt)
		{
		  default:
			o->wave = -1;
			o->mode = OSC_OFF;
			break;
		  case A2O_WNOISE:	o->mode = OSC_NOISE;	break;
		  case A2O_WWAVE:	o->mode = OSC_NOMIP;	break;
		  case A2O_WMIPWAVE:	o->mode = OSC_MIP;	break;
		}
		break;
	  }
	  case A2O_W_PITCH:
		ramp_set(&o->p, val + v->transpose + e->basepitch, start, dur);
		if(!dur)
			o->p_ramping = 1;
		break;
	  case A2O_W_AMP:
		ramp_set(&o->a, val, start, dur);
		break;
	  case A2O_W_PHASE:
		osc_set_phase(e, o, val, start);
		break;
	}
}

/* wtosc.c:108-126 */
static void osc_off(st_wtosc *o, int32_t *out, unsigned offset,
		unsigned frames, int add)
{
	ramp_prepare(&o->p, frames);
	ramp_prepare(&o->a, frames);
	ramp_run(&o->p, frames);
	ramp_run(&o->a, frames);
	if(!add)
		memset(out + offset, 0, frames * sizeof(int32_t));
}

/* wtosc.c:129-152 */
static void osc_noise(a2o_engine *e, st_wtosc *o, int32_t *out,
		unsigned offset, unsigned frames, int add)
{
	unsigned s, end = offset + frames;
	osc_run_pitch(o, frames);
	ramp_prepare(&o->a, frames);
	for(s = offset; s < end; ++s)
	{
		uint64_t nph = o->phase + o->dphase;
		int smp;
		if((o->dphase >= (1 << 23)) || ((nph ^ o->phase) >> 23))
			o->noise = a2o_noise(&e->noisestate) - 32767;
		o->phase = nph;
		smp = o->noise * (o->a.value >> 10) >> 6;
		if(add)
			out[s] += smp;
		else
			out[s] = smp;
		ramp_run(&o->a, 1);
	}
}

/* wtosc.c:200-236 */
static uint64_t osc_fragment(st_wtosc *o, const int16_t *d, int32_t *out,
		unsigned offset, unsigned frames, uint64_t ph, unsigned dph,
		int add, int looped, unsigned wsize)
{
	unsigned s, end = offset + frames;
	for(s = offset; s < end; ++s)
	{
		int v;
		int32_t smp;
		if(wsize)
		{
			if(looped)
				ph %= (uint64_t)wsize << 24;
			else if((ph >> 24) >= wsize)
			{
				if(!add)
					memset(out + s, 0, (end - s) *
							sizeof(int32_t));
				break;
			}
		}
		v = osc_inter(d, (unsigned)(ph >> 16), dph >> 16);
		smp = (int32_t)((int64_t)v * o->a.value >> 17);
		if(add)
			out[s] += smp;
		else
			out[s] = smp;
		ph += dph;
		ramp_run(&o->a, 1);
	}
	return ph;
}

/* wtosc.c:239-286; unloaded-wave check :168-183 */
static void osc_wavetable(a2o_engine *e, st_wtosc *o, int32_t *out,
		unsigned offset, unsigned frames, int add)
{
	unsigned mm, dph;
	uint64_t ph;
	wave *w = &e->waves[o->wave];
	if(!w->size[0])
	{
		o->wave = -1;
		o->mode = OSC_OFF;
		return;
	}
	osc_run_pitch(o, frames);
	dph = ((o->dphase + 255) >> 8) * w->period;
	ramp_prepare(&o->a, frames);
	for(mm = 0; (dph > (MAXPHINC << 8)) && (mm < A2O_MIPLEVELS - 1); ++mm)
		dph >>= 1;
	ph = o->phase >> mm;
	dph = (unsigned)((uint64_t)o->dphase * w->period >> mm);
	if(w->flags & A2O_LOOPED)
		ph %= (uint64_t)w->size[mm] << 24;
	else if((ph >> 24) > (w->size[mm] + A2O_WAVEPRE))
	{
		if(!add)
			memset(out + offset, 0, frames * sizeof(int32_t));
		return;
	}
	if(dph > (MAXPHINC << 16))
	{
		if(!add)
			memset(out + offset, 0, frames * sizeof(int32_t));
		ph += (uint64_t)dph * frames;
		o->phase = ph << mm;
		ramp_run(&o->a, frames);
	}
	else
		o->phase = osc_fragment(o, w->data[mm] + A2O_WAVEPRE, out,
				offset, frames, ph, dph, add, 0, 0) << mm;
}

/*
 * wtosc.c:301-358.  NOTE: the 32-bit 'size << 24' of :346 is restated as
 * written (unsigned wrap); it is a modulo by zero for looped non-mipmapped
 * waves whose size is a multiple of 256 - callers keep those out of tests,
 * like the reference's own data does.
 */
static void osc_wavetable_nomip(a2o_engine *e, st_wtosc *o, int32_t *out,
		unsigned offset, unsigned frames, int add)
{
	uint64_t dph;
	wave *w = &e->waves[o->wave];
	const int16_t *d = w->data[0] + A2O_WAVEPRE;
	if(!w->size[0])
	{
		o->wave = -1;
		o->mode = OSC_OFF;
		return;
	}
	osc_run_pitch(o, frames);
	dph = (uint64_t)o->dphase * w->period;
	ramp_prepare(&o->a, frames);
	if(dph >> 32)
	{
		if(!add)
			memset(out + offset, 0, frames * sizeof(int32_t));
		o->phase += dph * frames;
		ramp_run(&o->a, frames);
	}
	else if(dph > (MAXPHINC << 16))
	{
		o->phase = osc_fragment(o, d, out, offset, frames, o->phase,
				(unsigned)dph, add,
				(w->flags & A2O_LOOPED) ? 1 : 0, w->size[0]);
	}
	else
	{
		if(w->flags & A2O_LOOPED)
		{
			unsigned m = w->size[0] << 24;
			o->phase %= m;
		}
		else if((o->phase >> 24) > (w->size[0] + A2O_WAVEPRE))
		{
			if(!add)
				memset(out + offset, 0,
						frames * sizeof(int32_t));
			return;
		}
		o->phase = osc_fragment(o, d, out, offset, frames, o->phase,
				(unsigned)dph, add, 0, 0);
	}
}

static void osc_process(a2o_engine *e, st_wtosc *o, int32_t *out,
		unsigned offset, unsigned frames, int add)
{
	switch(o->mode)
	{
	  case OSC_OFF:
		osc_off(o, out, offset, frames, add);
		break;
	  case OSC_NOISE:
		osc_noise(e, o, out, offset, frames, add);
		break;
	  case OSC_NOMIP:
		osc_wavetable_nomip(e, o, out, offset, frames, add);
		break;
	  case OSC_MIP:
		osc_wavetable(e, o, out, offset, frames, add);
		break;
	}
}

/* ------------------------------------------------------------------ */
/* panmix.c                                                            */
/* ------------------------------------------------------------------ */

static void pm_init(st_panmix *pm)	/* panmix.c:252-262 */
{
	ramp_init(&pm->vol, 65536);
	ramp_init(&pm->pan, 0);
}

static void pm_write(st_panmix *pm, int reg, int val, unsigned start,
		unsigned dur)		/* panmix.c:287-295 */
{
	ramp_set(reg == A2O_PM_VOL ? &pm->vol : &pm->pan, val, start, dur);
}

/* panmix.c:117-122 etc: clamp variant is chosen per call */
static int pm_needs_clamp(const st_panmix *pm)
{
	return pm->pan.target > 0xffffff || pm->pan.target < -0xffffff ||
			pm->pan.value > 0xffffff || pm->pan.value < -0xffffff;
}

/* panmix.c:49-64 (1->1), :78-115 (1->2), :137-168 (2->1), :191-229 (2->2) */
static void pm_process(st_panmix *pm, int nin, int nout, int32_t **in,
		int32_t **out, unsigned offset, unsigned frames, int add)
{
	unsigned s, end = offset + frames;
	int clamp;
	if(nin == 1 && nout == 1)
	{
		ramp_prepare(&pm->vol, frames);
		for(s = offset; s < end; ++s)
		{
			int32_t r = (int32_t)((int64_t)in[0][s] *
					pm->vol.value >> 24);
			if(add)
				out[0][s] += r;
			else
				out[0][s] = r;
			ramp_run(&pm->vol, 1);
		}
		return;
	}
	clamp = pm_needs_clamp(pm);
	ramp_prepare(&pm->vol, frames);
	ramp_prepare(&pm->pan, frames);
	for(s = offset; s < end; ++s)
	{
		int vol = pm->vol.value;
		int vp = (int)((int64_t)pm->pan.value * vol >> 24);
		int v0 = vol - vp;
		int v1 = vol + vp;
		if(clamp)
		{
			if(v0 > vol << 1)
				v0 = vol << 1;
			if(v1 > vol << 1)
				v1 = vol << 1;
		}
		if(nin == 1)
		{
			int ins = in[0][s];
			int32_t r0 = (int32_t)((int64_t)ins * v0 >> 24);
			int32_t r1 = (int32_t)((int64_t)ins * v1 >> 24);
			if(add)
			{
				out[0][s] += r0;
				out[1][s] += r1;
			}
			else
			{
				out[0][s] = r0;
				out[1][s] = r1;
			}
		}
		else if(nout == 1)
		{
			int32_t r = (int32_t)(((int64_t)in[0][s] * v0 +
					(int64_t)in[1][s] * v1) >> 25);
			if(add)
				out[0][s] += r;
			else
				out[0][s] = r;
		}
		else
		{
			int in0 = in[0][s], in1 = in[1][s];
			int32_t r0 = (int32_t)((int64_t)in0 * v0 >> 24);
			int32_t r1 = (int32_t)((int64_t)in1 * v1 >> 24);
			if(add)
			{
				out[0][s] += r0;
				out[1][s] += r1;
			}
			else
			{
				out[0][s] = r0;
				out[1][s] = r1;
			}
		}
		ramp_run(&pm->vol, 1);
		ramp_run(&pm->pan, 1);
	}
}

/* ------------------------------------------------------------------ */
/* filter12.c                                                          */
/* ------------------------------------------------------------------ */

/* filter12.c:65-72: float multiply, then double sin() */
int a2o_f12_coeff(int cutoff_value, int samplerate)
{
	float f = a2o_p2i(cutoff_value >> 8) * (MIDDLEC / 16777216.0f);
	if(f > (samplerate >> 2))
		return 362 << 16;
	return (int)(512.0f * 65536.0f * sin(M_PI * f / samplerate));
}

/* The same for an array (test sweeps over the whole argument domain) */
void a2o_f12_coeff_array(const int *cutoff_values, int n, int samplerate, int *out)
{
	int i;
	for(i = 0; i < n; ++i)
		out[i] = a2o_f12_coeff(cutoff_values[i], samplerate);
}

/* filter12.c:141-177 */
static void f12_write(a2o_engine *e, voice *v, st_f12 *f, int reg, int val,
		unsigned start, unsigned dur)
{
	switch(reg)
	{
	  case A2O_F_CUTOFF:
		ramp_set(&f->cutoff, val + v->transpose, start, dur);
		if(dur < 256)
			f->f1 = a2o_f12_coeff(f->cutoff.value, e->samplerate);
		break;
	  case A2O_F_Q:
		if(val < 512)
			ramp_set(&f->q, 32768, start, dur);
		else
			ramp_set(&f->q, (65536 << 8) / val, start, dur);
		break;
	  case A2O_F_LP:	f->lp = val >> 8;	break;
	  case A2O_F_BP:	f->bp = val >> 8;	break;
	  case A2O_F_HP:	f->hp = val >> 8;	break;
	}
}

/* filter12.c:180-221 */
static void f12_init(a2o_engine *e, voice *v, st_f12 *f)
{
	ramp_init(&f->cutoff, 0);
	ramp_init(&f->q, 0);
	f12_write(e, v, f, A2O_F_CUTOFF, 0, 0, 0);
	f12_write(e, v, f, A2O_F_Q, 0, 0, 0);
	f->lp = 65536 >> 8;
	f->bp = f->hp = 0;
	f->d1[0] = f->d1[1] = f->d2[0] = f->d2[1] = 0;
}

/* filter12.c:74-119 */
static void f12_process(a2o_engine *e, st_f12 *f, int channels, int32_t **in,
		int32_t **out, unsigned offset, unsigned frames, int add)
{
	unsigned s, end = offset + frames;
	int c, df;
	int f0 = f->f1;
	ramp_prepare(&f->q, frames);
	ramp_prepare(&f->cutoff, frames);
	if(f->cutoff.delta)
	{
		ramp_run(&f->cutoff, frames);
		f->f1 = a2o_f12_coeff(f->cutoff.value, e->samplerate);
		df = (f->f1 - f0 + ((int)frames >> 1)) / (int)frames;
	}
	else
		df = 0;
	for(s = offset; s < end; ++s)
	{
		int fc = f0 >> 12;
		int q = f->q.value >> 12;
		for(c = 0; c < channels; ++c)
		{
			int d1 = f->d1[c] >> 4;
			int l = f->d2[c] + (fc * d1 >> 8);
			int h = (in[c][s] >> 5) - l - (q * d1 >> 8);
			int b = (fc * (h >> 4) >> 8) + f->d1[c];
			int fout = (l * f->lp + b * f->bp + h * f->hp) >> 3;
			if(add)
				out[c][s] += fout;
			else
				out[c][s] = fout;
			f->d1[c] = b;
			f->d2[c] = l;
		}
		f0 += df;
		ramp_run(&f->q, 1);
	}
}

/* ------------------------------------------------------------------ */
/* fm.c                                                                */
/* ------------------------------------------------------------------ */

/* fm.c:486-501: float argument, double sin(), float scale */
static void fm_build_sine(a2o_engine *e)
{
	int s;
	for(s = 0; s < 2049; ++s)
		e->fmsine[s] = sin(s * 2.0f * M_PI / 2048) * 32767.0f;
}

/* fm.c:111-122 */
static int32_t fm_osc(a2o_engine *e, st_fmop *o, int mod)
{
	int fb = (int)((int64_t)o->last * o->fb.value >> 17);
	unsigned ph = (o->phase + mod + fb) >> 5;
	o->last = a2o_lerp(e->fmsine, ph & ((2048 << 8) - 1));
	return (int32_t)((int64_t)o->last * o->a.value >> 16);
}

/* fm.c:125-140 */
static void fm_run_pitch(st_fmop *o, unsigned frames, int detune)
{
	int newpitch;
	ramp_prepare(&o->p, frames);
	ramp_run(&o->p, frames >> 1);
	newpitch = (o->p.value + detune) >> 8;
	if(newpitch != o->last_pitch)
	{
		o->dphase = a2o_p2i(newpitch);
		o->last_pitch = newpitch;
	}
}

/* fm.c:150-163 */
static int fm_sample(a2o_engine *e, st_fm *fm, int osbits)
{
	int i, v = 0;
	for(i = fm->nops - 1; i >= 0; --i)
	{
		if(i && fm->par == 1)
			v += fm_osc(e, &fm->op[i], 0);
		else
			v = fm_osc(e, &fm->op[i], v);
		fm->op[i].phase += fm->op[i].dphase >> osbits;
	}
	return v;
}

/* fm.c:170-192 */
static int fm_sample_rm(a2o_engine *e, st_fm *fm, int osbits)
{
	int i, v[2] = { 0, 0 };
	if(fm->nops == 2)
		for(i = 0; i < 2; ++i)
		{
			v[i] = fm_osc(e, &fm->op[i], 0);
			fm->op[i].phase += fm->op[i].dphase >> osbits;
		}
	else
		for(i = 0; i < 2; ++i)
		{
			v[i] = fm_osc(e, &fm->op[i],
					fm_osc(e, &fm->op[i + 2], 0));
			fm->op[i].phase += fm->op[i].dphase >> osbits;
			fm->op[i + 2].phase += fm->op[i + 2].dphase >> osbits;
		}
	return (int)((int64_t)v[0] * v[1] >> 23);
}

/* fm.c:194-233 */
static void fm_process(a2o_engine *e, st_fm *fm, int32_t *out,
		unsigned offset, unsigned frames, int add)
{
	int i, detune = 0;
	unsigned s, end = offset + frames;
	int osbits = fm->osbits;
	unsigned oversample = 1u << osbits;
	for(i = 0; i < fm->nops; ++i)
	{
		ramp_prepare(&fm->op[i].a, frames);
		ramp_prepare(&fm->op[i].fb, frames);
		fm_run_pitch(&fm->op[i], frames, detune);
		detune = fm->op[0].p.value;
	}
	for(s = offset; s < end; ++s)
	{
		unsigned os;
		int vsum = 0;
		for(os = 0; os < oversample; ++os)
			if(fm->par == 2)
				vsum += fm_sample_rm(e, fm, osbits);
			else
				vsum += fm_sample(e, fm, osbits);
		for(i = 0; i < fm->nops; ++i)
		{
			ramp_run(&fm->op[i].a, 1);
			ramp_run(&fm->op[i].fb, 1);
			fm->op[i].phase += fm->op[i].dphase & (oversample - 1);
		}
		if(add)
			out[s] += vsum >> osbits;
		else
			out[s] = vsum >> osbits;
	}
}

/* fm.c:328-336 */
static void fm_set_phase(st_fm *fm, int ph, unsigned sst)
{
	int i;
	for(i = 0; i < fm->nops; ++i)
	{
		int ssph = ph + (int)(sst * (fm->op[i].dphase >> 8) >> 8);
		fm->op[i].phase = (unsigned)(ssph * 2048 >> 8);
	}
}

/*
 * fm.c:339-408; oversampling per structure :236-321.
 *
 * AS BUILT: src/units/fm.c includes only fm.h -> a2_units.h and never sees
 * src/config.h, so its "#ifdef A2_HIFI" (fm.c:36) is false and the "normal"
 * table of fm.c:46-50 applies: fm1 1x, fm2/fm2r 2x, fm3/fm3p/fm4/fm4p/fm4r 4x
 * (fm4p and fm4r use A2FM3_OVERSAMPLE_BITS, fm.c:291-299, 313-321).  Verified
 * against the reference build: tests/golden fm_all, fm_bank64.
 */
static void fm_init(a2o_engine *e, voice *v, st_fm *fm, int kind,
		unsigned substart)
{
	int i;
	switch(kind)
	{
	  case A2O_FM1:	 fm->nops = 1; fm->osbits = 0; fm->par = 0; break;
	  case A2O_FM2:	 fm->nops = 2; fm->osbits = 1; fm->par = 0; break;
	  case A2O_FM3:	 fm->nops = 3; fm->osbits = 2; fm->par = 0; break;
	  case A2O_FM4:	 fm->nops = 4; fm->osbits = 2; fm->par = 0; break;
	  case A2O_FM3P: fm->nops = 3; fm->osbits = 2; fm->par = 1; break;
	  case A2O_FM4P: fm->nops = 4; fm->osbits = 2; fm->par = 1; break;
	  case A2O_FM2R: fm->nops = 2; fm->osbits = 1; fm->par = 2; break;
	  case A2O_FM4R: fm->nops = 4; fm->osbits = 2; fm->par = 2; break;
	}
	for(i = 0; i < fm->nops; ++i)
	{
		ramp_init(&fm->op[i].a, 0);
		ramp_init(&fm->op[i].fb, 0);
		ramp_init(&fm->op[i].p, v->transpose + e->basepitch);
		fm->op[i].last_pitch = 0;
		fm->op[i].last = 0;
	}
	fm->op[0].dphase = a2o_p2i(fm->op[0].p.value >> 8);
	for(i = 1; i < fm->nops; ++i)
		fm->op[i].dphase = fm->op[0].dphase;
	fm_set_phase(fm, 0, substart);
}

/* fm.c:411-483 */
static void fm_write(a2o_engine *e, voice *v, st_fm *fm, int reg, int val,
		unsigned start, unsigned dur)
{
	int op, which;
	if(reg == A2O_FM_PHASE)
	{
		fm_set_phase(fm, val, start);
		return;
	}
	op = (reg - 1) / 3;
	which = (reg - 1) % 3;
	if(op >= fm->nops)
		return;
	switch(which)
	{
	  case 0:
		if(!op)
			val += v->transpose + e->basepitch;
		ramp_set(&fm->op[op].p, val, start, dur);
		break;
	  case 1:
		ramp_set(&fm->op[op].a, val, start, dur);
		break;
	  case 2:
		ramp_set(&fm->op[op].fb, val, start, dur);
		break;
	}
}

/* ------------------------------------------------------------------ */
/* waveshaper.c                                                        */
/* ------------------------------------------------------------------ */

/* waveshaper.c:55-108 */
static void ws_process(st_ws *ws, int channels, int32_t **in, int32_t **out,
		unsigned offset, unsigned frames, int add)
{
	unsigned s, end = offset + frames;
	int c;
	ramp_prepare(&ws->amount, frames);
	for(s = offset; s < end; ++s)
	{
		int32_t a = ws->amount.value;
		int32_t a3p1 = (a << 1) + a + (1 << 24);
		int32_t asqr = (int32_t)((int64_t)(a >> 4) * (a >> 4) >> 24);
		for(c = 0; c < channels; ++c)
		{
			int32_t v = in[c][s];
			int32_t vsqr = (int32_t)((int64_t)v * v >> 22);
			int64_t vout = (int64_t)v * a3p1;
			int64_t sqrsub = (int64_t)a * vsqr;
			if(v >= 0)
				vout -= sqrsub;
			else
				vout += sqrsub;
			vout /= ((int64_t)asqr * vsqr >> 16) + (1 << 24);
			if(add)
				out[c][s] += (int32_t)vout;
			else
				out[c][s] = (int32_t)vout;
		}
		ramp_run(&ws->amount, 1);
	}
}

/* ------------------------------------------------------------------ */
/* Engine: waves, voices, groups                                       */
/* ------------------------------------------------------------------ */

a2o_engine *a2o_open(int samplerate, int channels)
{
	a2o_engine *e = (a2o_engine *)calloc(1, sizeof(a2o_engine));
	if(!e)
		return NULL;
	e->samplerate = samplerate;
	e->channels = channels < 2 ? 1 : 2;
	/* audiality2.c:398-399 with a2_F2Pf, pitch.c:45-48 */
	e->basepitch = (float)log2(MIDDLEC / (float)samplerate) * 65536.0f +
			0.5f;
	e->msdur = samplerate * 65.536f + .5f;	/* audiality2.c:499 */
	e->noisestate = 324357;		/* audiality2.h.cmake:62 */
	pm_init(&e->rootpm);
	fm_build_sine(e);
	if(!ptab_ready)
		ptab_build();
	return e;
}

void a2o_close(a2o_engine *e)
{
	int i, j;
	if(!e)
		return;
	for(i = 0; i < e->nwaves; ++i)
		for(j = 0; j < A2O_MIPLEVELS; ++j)
			free(e->waves[i].data[j]);
	free(e->waves);
	free(e->voices);
	{
		int gi;
		for(gi = 0; gi < e->ngroups; ++gi)
			if(e->groups[gi].has_fbd)
			{
				free(e->groups[gi].fbd.lbuf);
				free(e->groups[gi].fbd.rbuf);
			}
	}
	free(e->groups);
	free(e->order);
	free(e);
}

int a2o_basepitch(a2o_engine *e)	{ return e->basepitch; }
uint32_t a2o_msdur(a2o_engine *e)	{ return e->msdur; }
void a2o_set_noiseseed(a2o_engine *e, uint32_t seed) { e->noisestate = seed; }

int a2o_upload_wave(a2o_engine *e, int type, unsigned period, unsigned flags,
		const int16_t *data, unsigned length)
{
	wave *w;
	wave *nw = (wave *)realloc(e->waves, sizeof(wave) * (e->nwaves + 1));
	if(!nw)
		return -1;
	e->waves = nw;
	w = &e->waves[e->nwaves];
	memset(w, 0, sizeof(wave));
	w->type = type;
	w->flags = flags;
	w->period = period;
	if(type == A2O_WWAVE || type == A2O_WMIPWAVE)
		if(wave_prepare(w, data, length))
			return -1;
	return e->nwaves++;
}

/* waves.c:629-708 */
int a2o_builtin_wave(a2o_engine *e, const char *name)
{
	int i, s, h = -1;
	int16_t buf[WAVEPERIOD];
	for(i = 0; i < e->nwaves; ++i)
		if(!strcmp(e->waves[i].name, name))
			return i;
	if(!strcmp(name, "off"))
		h = a2o_upload_wave(e, A2O_WOFF, 0, 0, NULL, 0);
	else if(!strcmp(name, "noise"))
		h = a2o_upload_wave(e, A2O_WNOISE, 256, A2O_LOOPED, NULL, 0);
	else
	{
		int duty = 0;
		if(!strcmp(name, "square"))
			duty = 50;
		else if(!strncmp(name, "pulse", 5))
			duty = atoi(name + 5);
		if(duty)
		{
			int s1 = (WAVEPERIOD * duty + 50) / 100;
			memset(buf, 0, sizeof(buf));
			/*
			 * waves.c:641-644 - the sample at index s1 is skipped
			 * by the '++s' and keeps what the previous, narrower
			 * duty cycle left there: -32767 for every duty > 1
			 * (uninitialised stack for pulse1; we use the same).
			 */
			for(s = 0; s < s1; ++s)
				buf[s] = 32767;
			for(s = s1; s < WAVEPERIOD; ++s)
				buf[s] = -32767;
		}
		else if(!strcmp(name, "saw"))
			for(s = 0; s < WAVEPERIOD; ++s)
				buf[s] = s * 65534 / WAVEPERIOD - 32767;
		else if(!strcmp(name, "triangle"))
		{
			/* rendered over the saw left in 'buf', waves.c:655-663 */
			for(s = 0; s < WAVEPERIOD; ++s)
				buf[s] = s * 65534 / WAVEPERIOD - 32767;
			for(s = 0; s < WAVEPERIOD / 2; ++s)
				buf[(5 * WAVEPERIOD / 4 - s - 1) % WAVEPERIOD] =
					buf[s + WAVEPERIOD / 4] =
					s * 65534 * 2 / WAVEPERIOD - 32767;
		}
		else if(!strcmp(name, "sine") || !strcmp(name, "asine") ||
				!strcmp(name, "hsine") || !strcmp(name, "qsine"))
		{
			for(s = 0; s < WAVEPERIOD; ++s)
				buf[s] = sin(s * 2.0f * M_PI / WAVEPERIOD) *
						32767.0f;
			if(strcmp(name, "sine"))
				for(s = WAVEPERIOD / 2; s < WAVEPERIOD; ++s)
					buf[s] = -buf[s];
			if(!strcmp(name, "hsine") || !strcmp(name, "qsine"))
				for(s = WAVEPERIOD / 2; s < WAVEPERIOD; ++s)
					buf[s] = 0;
			if(!strcmp(name, "qsine"))
				for(s = 0; s < WAVEPERIOD / 4; ++s)
					buf[s + WAVEPERIOD / 2] = buf[s];
		}
		else
			return -1;
		h = a2o_upload_wave(e, A2O_WMIPWAVE, WAVEPERIOD, A2O_LOOPED,
				buf, WAVEPERIOD);
	}
	if(h >= 0)
	{
		strncpy(e->waves[h].name, name, sizeof(e->waves[h].name) - 1);
	}
	return h;
}

const int16_t *a2o_wave_data(a2o_engine *e, int wv, int level, unsigned *size)
{
	if(wv < 0 || wv >= e->nwaves || level < 0 || level >= A2O_MIPLEVELS)
		return NULL;
	if(size)
		*size = e->waves[wv].size[level];
	return e->waves[wv].data[level];
}

static int order_push(a2o_engine *e, int item)
{
	if(e->norder == e->corder)
	{
		int nc = e->corder ? e->corder * 2 : 256;
		int *no = (int *)realloc(e->order, sizeof(int) * nc);
		if(!no)
			return -1;
		e->order = no;
		e->corder = nc;
	}
	e->order[e->norder++] = item;
	return 0;
}

int a2o_new_group(a2o_engine *e)
{
	group *ng = (group *)realloc(e->groups, sizeof(group) *
			(e->ngroups + 1));
	if(!ng)
		return -1;
	e->groups = ng;
	memset(&e->groups[e->ngroups], 0, sizeof(group));
	pm_init(&e->groups[e->ngroups].pm);
	if(order_push(e, ~e->ngroups))
		return -1;
	return e->ngroups++;
}

/*
 * Put a `fbdelay` between the group's inline and its panmix, i.e. the song-level chain
 * { inline 0 *; fbdelay * *; panmix * > } (benchmark/k2trance.a2s:920-924), and write its seven
 * registers (fbdelay, ldelay, rdelay in ms; drygain, fbgain, lgain, rgain; all 16:16).
 */
int a2o_group_fbdelay(a2o_engine *e, int group, const int32_t *regs)
{
	int r;
	st_fbdelay *f;
	if(!e || group < 0 || group >= e->ngroups || !regs)
		return -1;
	f = &e->groups[group].fbd;
	if(!e->groups[group].has_fbd && fbd_init(f, e->samplerate))
		return -1;
	e->groups[group].has_fbd = 1;
	for(r = 0; r < 7; ++r)
		fbd_write(f, r, regs[r], e->samplerate);
	return 0;
}

int a2o_new_voice(a2o_engine *e, const a2o_unitspec *chain, int nunits,
		int transpose, unsigned substart, int grp)
{
	voice *v;
	int i;
	if(nunits < 1 || nunits > A2O_MAXUNITS)
		return -1;
	if(e->nvoices == e->cvoices)
	{
		int nc = e->cvoices ? e->cvoices * 2 : 256;
		voice *nv = (voice *)realloc(e->voices, sizeof(voice) * nc);
		if(!nv)
			return -1;
		e->voices = nv;
		e->cvoices = nc;
	}
	v = &e->voices[e->nvoices];
	memset(v, 0, sizeof(voice));
	v->alive = 1;
	v->nunits = nunits;
	v->transpose = transpose;
	v->group = grp;
	for(i = 0; i < nunits; ++i)
	{
		unit *u = &v->u[i];
		u->spec = chain[i];
		switch(chain[i].kind)
		{
		  case A2O_WTOSC:
			osc_init(e, v, &u->s.osc, substart);
			break;
		  case A2O_PANMIX:
			pm_init(&u->s.pm);
			break;
		  case A2O_FILTER12:
			f12_init(e, v, &u->s.f12);
			break;
		  case A2O_WAVESHAPER:
			ramp_init(&u->s.ws.amount, 0);
			break;
		  default:
			if(chain[i].kind < A2O_FM1 || chain[i].kind > A2O_FM4R)
				return -1;
			fm_init(e, v, &u->s.fm, chain[i].kind, substart);
			break;
		}
	}
	if(grp < 0 && order_push(e, e->nvoices))
		return -1;
	return e->nvoices++;
}

void a2o_kill_voice(a2o_engine *e, int vi)
{
	if(vi >= 0 && vi < e->nvoices)
		e->voices[vi].alive = 0;
}

void a2o_write(a2o_engine *e, int vi, int ui, int reg, int value,
		unsigned start, unsigned dur)
{
	voice *v = &e->voices[vi];
	unit *u = &v->u[ui];
	switch(u->spec.kind)
	{
	  case A2O_WTOSC:
		osc_write(e, v, &u->s.osc, reg, value, start, dur);
		break;
	  case A2O_PANMIX:
		pm_write(&u->s.pm, reg, value, start, dur);
		break;
	  case A2O_FILTER12:
		f12_write(e, v, &u->s.f12, reg, value, start, dur);
		break;
	  case A2O_WAVESHAPER:
		ramp_set(&u->s.ws.amount, value, start, dur);
		break;
	  default:
		fm_write(e, v, &u->s.fm, reg, value, start, dur);
		break;
	}
}

/* Bulk form of a2o_write for large banks: voice first + i gets values[i * stride] */
void a2o_write_all(a2o_engine *e, int first, int count, int ui, int reg,
		const int *values, int stride, unsigned start, unsigned dur)
{
	int i;
	for(i = 0; i < count; ++i)
		a2o_write(e, first + i, ui, reg, values[(long)i * stride],
				start, dur);
}

/* ------------------------------------------------------------------ */
/* Segment loop and bus semantics (core.c:1847-1896, 1749-1776)        */
/* ------------------------------------------------------------------ */

/* One round of "for(u = v->units; u; u = u->next) u->Process(u, s, res)" */
static void voice_run_units(a2o_engine *e, voice *v, int32_t **bus,
		unsigned offset, unsigned frames)
{
	int i;
	int32_t *scr[2] = { e->vscratch[0], e->vscratch[1] };
	for(i = 0; i < v->nunits; ++i)
	{
		unit *u = &v->u[i];
		int32_t **out = u->spec.wireout ? bus : scr;
		switch(u->spec.kind)
		{
		  case A2O_WTOSC:
			osc_process(e, &u->s.osc, out[0], offset, frames,
					u->spec.add);
			break;
		  case A2O_PANMIX:
			pm_process(&u->s.pm, u->spec.ninputs, u->spec.noutputs,
					scr, out, offset, frames, u->spec.add);
			break;
		  case A2O_FILTER12:
			f12_process(e, &u->s.f12, u->spec.ninputs, scr, out,
					offset, frames, u->spec.add);
			break;
		  case A2O_WAVESHAPER:
			ws_process(&u->s.ws, u->spec.ninputs, scr, out,
					offset, frames, u->spec.add);
			break;
		  default:
			fm_process(e, &u->s.fm, out[0], offset, frames,
					u->spec.add);
			break;
		}
	}
}

/*
 * core.c:1784-1835 + 1847-1880 for a voice whose "VM" is the event list:
 * at frame s, everything due within that frame (time - now <= 255) is applied
 * with start = time & 255 (core.c:143-149), then the units run up to the next
 * wake-up or the end of the enclosing segment.
 */
static void run_events(a2o_engine *e, int *evi, int nev, int *evpos,
		unsigned now, voice *v, int vi, st_panmix *pm)
{
	while(*evpos < nev)
	{
		const a2o_event *ev = &e->ev[evi[*evpos]];
		if((int32_t)(ev->time - now) > 255)
			break;
		if(ev->kind == A2O_EV_WRITE && v)
			a2o_write(e, vi, ev->unit, ev->reg, ev->value,
					ev->time & 255, ev->dur);
		else if(ev->kind != A2O_EV_WAKE && pm && ev->reg >= 0)
			pm_write(pm, ev->reg, ev->value, ev->time & 255,
					ev->dur);
		++*evpos;
	}
}

static unsigned next_wake(a2o_engine *e, int *evi, int nev, int evpos,
		unsigned now, unsigned maxframes)
{
	if(evpos < nev)
	{
		unsigned d = (e->ev[evi[evpos]].time - now) >> 8;
		if(d < maxframes)
			return d;
	}
	return maxframes;
}

static void voice_process(a2o_engine *e, int vi, int32_t **bus,
		unsigned offset, unsigned frames)
{
	voice *v = &e->voices[vi];
	unsigned s = offset, stop = offset + frames;
	if(!v->alive)
		return;
	while(s < stop)
	{
		unsigned now = e->now_fragstart + (s << 8);
		unsigned res;
		run_events(e, v->evi, v->nev, &v->evpos, now, v, vi, NULL);
		res = next_wake(e, v->evi, v->nev, v->evpos, now, stop - s);
		voice_run_units(e, v, bus, s, res);
		s += res;
	}
}

/* fbdelay.c:176-208: zeroed delay lines, default registers */
static int fbd_init(st_fbdelay *f, int samplerate)
{
	memset(f, 0, sizeof(*f));
	f->lbuf = (int32_t *)calloc(A2O_FBD_BUFSIZE, sizeof(int32_t));
	f->rbuf = (int32_t *)calloc(A2O_FBD_BUFSIZE, sizeof(int32_t));
	if(!f->lbuf || !f->rbuf)
		return -1;
	f->fbdelay = (int)((int64_t)(400 << 16) * samplerate / 65536000);
	f->ldelay = (int)((int64_t)(280 << 16) * samplerate / 65536000);
	f->rdelay = (int)((int64_t)(320 << 16) * samplerate / 65536000);
	f->drygain = 65536;
	f->fbgain = 16384;
	f->lgain = 32768;
	f->rgain = 32768;
	return 0;
}

/* fbdelay.c:229-270: register write callbacks (start / duration are ignored there) */
static void fbd_write(st_fbdelay *f, int reg, int v, int samplerate)
{
	switch(reg)
	{
	  case 0: f->fbdelay = (int)((int64_t)v * samplerate / 65536000); break;
	  case 1: f->ldelay = (int)((int64_t)v * samplerate / 65536000); break;
	  case 2: f->rdelay = (int)((int64_t)v * samplerate / 65536000); break;
	  case 3: f->drygain = v; break;
	  case 4: f->fbgain = v; break;
	  case 5: f->lgain = v; break;
	  case 6: f->rgain = v; break;
	}
}

/* fbdelay.c:68-127, the 2 -> 2 replacing variant (fbdelay_Process22), in place */
static void fbd_process22(st_fbdelay *f, int32_t **buf, unsigned offset,
		unsigned frames)
{
	unsigned s, end = offset + frames;
#define	WI(x)	((f->bufpos - (x)) & (A2O_FBD_BUFSIZE - 1))
	for(s = offset; s < end; ++s)
	{
		int i0 = buf[0][s];
		int i1 = buf[1][s];
		/* feedback taps, cross-fed ("reverse stereo") */
		int o0 = (int)((int64_t)f->rbuf[WI(f->fbdelay)] * f->fbgain >> 16);
		int o1 = (int)((int64_t)f->lbuf[WI(f->fbdelay)] * f->fbgain >> 16);
		f->lbuf[WI(0)] = i0 + o0;
		f->rbuf[WI(0)] = i1 + o1;
		o0 += (int)((int64_t)f->lbuf[WI(f->ldelay)] * f->lgain >> 16);
		o1 += (int)((int64_t)f->rbuf[WI(f->rdelay)] * f->rgain >> 16);
		o0 += (int)((int64_t)i0 * f->drygain >> 16);
		o1 += (int)((int64_t)i1 * f->drygain >> 16);
		buf[0][s] = o0;
		buf[1][s] = o1;
		++f->bufpos;
	}
#undef	WI
}

/* group = { inline 0 *; [fbdelay * *;] panmix * *; xinsert * > } (audiality2.c:294-304) */
static void group_process(a2o_engine *e, int gi, unsigned offset,
		unsigned frames)
{
	group *g = &e->groups[gi];
	unsigned s = offset, stop = offset + frames;
	int32_t *gb[2] = { e->groupbus[0], e->groupbus[1] };
	int32_t *rb[2] = { e->rootbus[0], e->rootbus[1] };
	while(s < stop)
	{
		unsigned now = e->now_fragstart + (s << 8);
		unsigned res, c, j;
		int vi;
		run_events(e, g->evi, g->nev, &g->evpos, now, NULL, 0, &g->pm);
		res = next_wake(e, g->evi, g->nev, g->evpos, now, stop - s);
		/* inline, replacing: core.c:1769-1776 */
		for(c = 0; c < 2; ++c)
			memset(gb[c] + s, 0, res * sizeof(int32_t));
		/* newest voice first: a2_VoiceNew links at the head, :476-477 */
		for(vi = e->nvoices - 1; vi >= 0; --vi)
			if(e->voices[vi].group == gi)
				voice_process(e, vi, gb, s, res);
		if(g->has_fbd)
			fbd_process22(&g->fbd, gb, s, res);
		pm_process(&g->pm, 2, 2, gb, gb, s, res, 0);
		/* xinsert without clients: bypass-add, xinsert.c:149-156 */
		for(c = 0; c < 2; ++c)
			for(j = s; j < s + res; ++j)
				rb[c][j] += gb[c][j];
		s += res;
	}
}

/* root = { inline 0 *|2; panmix * *|2 1; xinsert * > } (audiality2.c:268-292) */
static void root_process(a2o_engine *e, unsigned frames)
{
	unsigned s = 0, stop = frames;
	int32_t *rb[2] = { e->rootbus[0], e->rootbus[1] };
	int32_t *mb[2] = { e->master[0], e->master[1] };
	while(s < stop)
	{
		unsigned now = e->now_fragstart + (s << 8);
		unsigned res, c, j;
		int k;
		run_events(e, e->rootevi, e->rootnev, &e->rootevpos, now,
				NULL, 0, &e->rootpm);
		res = next_wake(e, e->rootevi, e->rootnev, e->rootevpos, now,
				stop - s);
		for(c = 0; c < 2; ++c)
			memset(rb[c] + s, 0, res * sizeof(int32_t));
		for(k = e->norder - 1; k >= 0; --k)
		{
			if(e->order[k] >= 0)
				voice_process(e, e->order[k], rb, s, res);
			else
				group_process(e, ~e->order[k], s, res);
		}
		pm_process(&e->rootpm, 2, e->channels, rb, rb, s, res, 0);
		for(c = 0; c < (unsigned)e->channels; ++c)
			for(j = s; j < s + res; ++j)
				mb[c][j] += rb[c][j];
		s += res;
	}
}

static int *collect(const a2o_event *ev, int nev, int kind_lo, int kind_hi,
		int target, int *count)
{
	int i, n = 0;
	int *idx;
	for(i = 0; i < nev; ++i)
		if(ev[i].kind >= kind_lo && ev[i].kind <= kind_hi &&
				ev[i].voice == target)
			++n;
	idx = (int *)malloc(sizeof(int) * (n ? n : 1));
	n = 0;
	for(i = 0; i < nev; ++i)
		if(ev[i].kind >= kind_lo && ev[i].kind <= kind_hi &&
				ev[i].voice == target)
			idx[n++] = i;
	*count = n;
	return idx;
}

/* core.c:1927-2001 (fragment loop), bufferdrv.c:28-40 (driver buffer) */
void a2o_render(a2o_engine *e, const a2o_event *ev, int nev,
		int32_t *out, long frames, int buffer)
{
	long done = 0;
	int i, *cnt, *pos, *flat;
	e->ev = ev;
	/* per-voice event index lists (stable) */
	cnt = (int *)calloc(e->nvoices + 1, sizeof(int));
	pos = (int *)calloc(e->nvoices + 1, sizeof(int));
	flat = (int *)malloc(sizeof(int) * (nev ? nev : 1));
	for(i = 0; i < nev; ++i)
		if(ev[i].kind <= A2O_EV_WAKE && ev[i].voice >= 0 &&
				ev[i].voice < e->nvoices)
			++cnt[ev[i].voice];
	for(i = 1; i <= e->nvoices; ++i)
		pos[i] = pos[i - 1] + cnt[i - 1];
	for(i = 0; i < e->nvoices; ++i)
	{
		e->voices[i].evi = flat + pos[i];
		e->voices[i].nev = 0;
		e->voices[i].evpos = 0;
	}
	for(i = 0; i < nev; ++i)
		if(ev[i].kind <= A2O_EV_WAKE && ev[i].voice >= 0 &&
				ev[i].voice < e->nvoices)
		{
			voice *v = &e->voices[ev[i].voice];
			v->evi[v->nev++] = i;
		}
	for(i = 0; i < e->ngroups; ++i)
	{
		e->groups[i].evi = collect(ev, nev, A2O_EV_GROUPWRITE,
				A2O_EV_GROUPWRITE, i, &e->groups[i].nev);
		e->groups[i].evpos = 0;
	}
	e->rootevi = collect(ev, nev, A2O_EV_ROOTWRITE, A2O_EV_ROOTWRITE, 0,
			&e->rootnev);
	e->rootevpos = 0;

	while(done < frames)
	{
		long n = frames - done < buffer ? frames - done : buffer;
		long off = 0;
		while(off < n)
		{
			unsigned frag = n - off > A2O_MAXFRAG ? A2O_MAXFRAG :
					(unsigned)(n - off);
			unsigned c, s;
			memset(e->master, 0, sizeof(e->master));
			root_process(e, frag);
			for(s = 0; s < frag; ++s)
				for(c = 0; c < (unsigned)e->channels; ++c)
					out[(done + off + s) * e->channels + c] =
							e->master[c][s];
			off += frag;
			e->now_fragstart += frag << 8;
		}
		done += n;
	}
	for(i = 0; i < e->ngroups; ++i)
		free(e->groups[i].evi);
	free(e->rootevi);
	free(flat);
	free(cnt);
	free(pos);
}
