/*
 * a2render.c - offline render harness around the Audiality 2 public API.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/): this program is linked against
 *   - oracle/_ref/libaudiality2.so          -> "a2render"      (the oracle)
 *   - oracle/_ref/libaudiality2_host.so +
 *     audiality2_b200/liba2cu_units.so      -> "a2render_cuda" (drop-in check)
 * and dumps the int32 8:24 master output so the two can be compared bit for
 * bit. It uses only the public surface (a2_OpenConfig / a2_NewDriver /
 * a2_Open / a2_Load / a2_Get / a2_Starta / a2_Run, include/audiality2.h.cmake
 * :141-424, include/a2_drivers.h:93-139) exactly like a2play/a2play.c:664-765
 * and src/render.c:34-127 do, with the "buffer" audio driver
 * (src/drivers/bufferdrv.c:28-40) or, with -DA2CU_PLUGIN, the "cuda" driver
 * registered by the plug-in.
 *
 * Output: raw little-endian int32, interleaved [frame][channel].
 * A one-line JSON summary (frames, seconds spent inside the a2_Run loop, RT
 * error, active voices) goes to stdout.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "audiality2.h"
#include "a2_waves.h"

#ifdef A2CU_PLUGIN
extern int a2cu_RegisterDriver(void);
#endif

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static void die(const char *what, int err)
{
	fprintf(stderr, "a2render: %s (%s)\n", what, a2_ErrorString(err));
	exit(1);
}

int main(int argc, char **argv)
{
	int rate = 48000, buffer = 64, channels = 2, nargs = 0, i, copies = 1;
	int shard_rank = 0, shard_world = 1;
	long frames = 48000, warmup = 0, done = 0;
	int args[16];
	int noiseseed = -1;
	const char *prog = "Song", *out = NULL, *file = NULL;
	const char *driver = "buffer";
	const char *dumpwave = NULL;
	const char *upload = NULL;
	A2_config *cfg;
	A2_driver *drv;
	A2_interface *iface;
	A2_handle bank, ph, vh;
	A2_audiodriver *ad;
	FILE *f = NULL;
	double t0, t = 0.0;
	int32_t *il;
	int voices = 0, rterr;

	for(i = 1; i < argc; ++i)
	{
		if(!strcmp(argv[i], "-r")) rate = atoi(argv[++i]);
		else if(!strcmp(argv[i], "-b")) buffer = atoi(argv[++i]);
		else if(!strcmp(argv[i], "-c")) channels = atoi(argv[++i]);
		else if(!strcmp(argv[i], "-n")) frames = atol(argv[++i]);
		else if(!strcmp(argv[i], "-w")) warmup = atol(argv[++i]);
		else if(!strcmp(argv[i], "-p")) prog = argv[++i];
		else if(!strcmp(argv[i], "-o")) out = argv[++i];
		else if(!strcmp(argv[i], "-d")) driver = argv[++i];
		else if(!strcmp(argv[i], "-s")) noiseseed = atoi(argv[++i]);
		else if(!strcmp(argv[i], "-W")) dumpwave = argv[++i];
		else if(!strcmp(argv[i], "-x")) copies = atoi(argv[++i]);
		else if(!strcmp(argv[i], "-X"))		/* -X rank/world: start only copies k with k % world == rank */
		{
			if(sscanf(argv[++i], "%d/%d", &shard_rank, &shard_world) != 2 ||
					shard_world < 1 || shard_rank < 0 ||
					shard_rank >= shard_world)
			{
				fprintf(stderr, "a2render: -X rank/world\n");
				return 2;
			}
		}
		else if(!strcmp(argv[i], "-U")) upload = argv[++i];
		else if(!strcmp(argv[i], "-a"))
		{
			/* 16:16 fixed point, same conversion as a2_Start() */
			if(nargs < 16)
				args[nargs++] = (int)(atof(argv[++i]) * 65536.0);
		}
		else file = argv[i];
	}
	if(!file)
	{
		fprintf(stderr, "usage: a2render [-r rate] [-b buffer] "
				"[-c channels] [-n frames] [-w warmupframes] "
				"[-p program] [-a arg]... [-s noiseseed] "
				"[-d driver] [-o out.raw] file.a2s\n");
		return 2;
	}
#ifdef A2CU_PLUGIN
	if(a2cu_RegisterDriver())
	{
		fprintf(stderr, "a2render: could not register cuda driver\n");
		return 1;
	}
#endif
	if(!(cfg = a2_OpenConfig(rate, buffer, channels, A2_AUTOCLOSE)))
		die("a2_OpenConfig", a2_LastError());
	if(!(drv = a2_NewDriver(A2_AUDIODRIVER, driver)))
		die("a2_NewDriver", a2_LastError());
	if(a2_AddDriver(cfg, drv))
		die("a2_AddDriver", a2_LastError());
	if(!(iface = a2_Open(cfg)))
		die("a2_Open", a2_LastError());
	if(noiseseed >= 0)
		a2_SetStateProperty(iface, A2_PNOISESEED, noiseseed);
	if(dumpwave)
	{
		/*
		 * Checksum + the samples around the duty-cycle edge of a builtin
		 * wave: a2_InitWaves() never writes buf[s1] of "pulse1"
		 * (src/waves.c:639-646, `for(++s; ...)` skips one sample; later
		 * pulses inherit -32767 from the previous one), so that sample is
		 * whatever the stack held - it differs between processes.
		 */
		A2_handle wh = a2_Get(iface, A2_ROOTBANK, dumpwave);
		A2_wave *w = wh >= 0 ? a2_GetWave(iface, wh) : NULL;
		if(w && (w->type == A2_WWAVE || w->type == A2_WMIPWAVE))
		{
			unsigned k, sum = 0;
			for(k = 0; k < w->d.wave.size[0]; ++k)
				sum = sum * 31 + (unsigned short)
						w->d.wave.data[0][A2_WAVEPRE + k];
			fprintf(stderr, "wave %s size %u sum %u s[18..22] %d %d %d %d %d\n",
					dumpwave, w->d.wave.size[0], sum,
					w->d.wave.data[0][A2_WAVEPRE + 18],
					w->d.wave.data[0][A2_WAVEPRE + 19],
					w->d.wave.data[0][A2_WAVEPRE + 20],
					w->d.wave.data[0][A2_WAVEPRE + 21],
					w->d.wave.data[0][A2_WAVEPRE + 22]);
		}
		else
			fprintf(stderr, "wave %s: not a table\n", dumpwave);
	}
	if((bank = a2_Load(iface, file, 0)) < 0)
		die("a2_Load", -bank);
	if((ph = a2_Get(iface, bank, prog)) < 0)
		die("a2_Get(program)", -ph);
	if(upload)
	{
		/*
		 * -U type:period:flags:length:seed  uploads a pseudo-random int16
		 * wave through the public API (a2_UploadWave, a2_waves.h:148) and
		 * passes its handle as the program's LAST argument, so scripts can
		 * play sampled waves: `w W` with W the handle. The generator is the
		 * LCG tests/test_sampled_waves.py repeats.
		 */
		int wt = 0, wperiod = 0, wflags = 0, wlen = 0, k;
		unsigned wseed = 1;
		int16_t *wd;
		A2_handle wh;
		if(sscanf(upload, "%d:%d:%d:%d:%u", &wt, &wperiod, &wflags, &wlen,
				&wseed) != 5 || wlen < 1)
		{
			fprintf(stderr, "a2render: bad -U spec\n");
			return 2;
		}
		wd = (int16_t *)malloc(sizeof(int16_t) * wlen);
		for(k = 0; k < wlen; ++k)
		{
			wseed = wseed * 1664525u + 1013904223u;
			wd[k] = (int16_t)((int)(wseed >> 16) % 50001 - 25000);
		}
		wh = a2_UploadWave(iface, (A2_wavetypes)wt, wperiod, wflags, A2_I16,
				wd, sizeof(int16_t) * wlen);
		free(wd);
		if(wh < 0)
			die("a2_UploadWave", -wh);
		if(nargs < 16)
			args[nargs++] = wh << 16;
	}
	a2_TimestampReset(iface);
	/*
	 * -x N: start the program N times under the root voice (BASELINE config
	 * 5's "voice-spawn stress multiplier"); copy k gets its first argument
	 * (P, transpose) offset by (k % 25 - 12) semitones.
	 */
	for(i = 0; i < copies; ++i)
	{
		int a[16];
		int na = nargs;
		/*
		 * Sharding over engine states / GPUs: whole sub-trees
		 * under the root, dealt round-robin; the int32 outputs of the shards are
		 * summed by the caller.
		 */
		if(i % shard_world != shard_rank)
			continue;
		memcpy(a, args, sizeof(a));
		if(copies > 1)
		{
			if(!na)
			{
				a[0] = 0;
				na = 1;
			}
			a[0] += ((i % 25) - 12) * 65536 / 12;
		}
		if((vh = a2_Starta(iface, a2_RootVoice(iface), ph, na, a)) < 0)
			die("a2_Starta", -vh);
	}

	ad = (A2_audiodriver *)drv;
	if(out && !(f = fopen(out, "wb")))
	{
		perror(out);
		return 1;
	}
	il = (int32_t *)malloc(sizeof(int32_t) * buffer * channels);
	while(done < warmup + frames)
	{
		unsigned n = buffer;
		int c, s, res;
		if(done < warmup && n > warmup - done)
			n = warmup - done;
		else if(n > warmup + frames - done)
			n = warmup + frames - done;
		t0 = now_s();
		if((res = a2_Run(iface, n)) < 0)
			die("a2_Run", -res);
		if(done >= warmup)
			t += now_s() - t0;
		a2_PumpMessages(iface);
		if(f && done >= warmup)
		{
			for(s = 0; s < (int)n; ++s)
				for(c = 0; c < channels; ++c)
					il[s * channels + c] =
							ad->buffers[c][s];
			fwrite(il, sizeof(int32_t), n * channels, f);
		}
		done += n;
	}
	a2_GetStateProperty(iface, A2_PACTIVEVOICES, &voices);
	rterr = a2_LastRTError(iface);
	printf("{\"frames\": %ld, \"seconds\": %.6f, \"channels\": %d, "
			"\"rate\": %d, \"buffer\": %d, \"active_voices\": %d, "
			"\"rt_error\": %d}\n",
			frames, t, channels, rate, buffer, voices, rterr);
	if(f)
		fclose(f);
	free(il);
	a2_Close(iface);
	return 0;
}
