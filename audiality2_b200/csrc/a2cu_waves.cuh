// a2cu_waves.cuh - wave preparation on the device.
//
// The reference prepares a wave on the CPU when it is uploaded or rendered: pad samples around
// every level (src/waves.c:90-106), nine mip levels by a (1 2 1)/4 decimation (:108-151). Here only
// the raw int16 samples cross the bus; these kernels fill in the rest inside the device pool, and
// additionally tabulate the two-stage Hermite coefficients (a2_Hermite2c, include/a2_dsp.h:83-89)
// that the oscillator kernels gather. Bit-exact integer arithmetic; compared with the port's
// prepared data in tests/test_cuda_parity.py and end to end by every golden case.
#pragma once
#include "a2cu_device.cuh"

namespace a2cu {

constexpr int kWavePost = 132;      // A2_WAVEPOST, a2_waves.h:63-64

// lvl points at the level's first pad sample; n samples follow, then kWavePost pad samples.
__global__ void wave_pad(int16_t *lvl, unsigned n, int looped) {
    int16_t *d = lvl + kWavePre;
    const bool loop = looped && n;
    if (threadIdx.x == 0) lvl[0] = loop ? d[n - 1] : (int16_t)0;          // pre pad = last sample
    for (int i = threadIdx.x; i < kWavePost; i += blockDim.x) d[n + i] = loop ? d[i % n] : (int16_t)0;
}

// waves.c:108-151: d[s] = (2 * sd[2s] + sd[2s - 1] + sd[2s + 1]) >> 2 (sd padded on both sides)
__global__ void wave_mip(const int16_t *sd, int16_t *d, unsigned n) {
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) d[s] = (int16_t)((((int)sd[s * 2] << 1) + sd[(long long)s * 2 - 1] + sd[s * 2 + 1]) >> 2);
}

// a2_dsp.h:83-89 for every position the oscillator can address: {d0, a, b, c}
__global__ void wave_coef(const int16_t *d, int4 *out, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int dm = d[k - 1], d0 = d[k], d1 = d[k + 1], d2 = d[k + 2];
    const int c = (d1 - dm) >> 1;
    const int a = (3 * (d0 - d1) + d2 - dm) >> 1;
    const int b = dm - d0 + c - a;
    out[k] = make_int4(d0, a, b, c);
}

}  // namespace a2cu
