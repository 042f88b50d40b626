// a2cu_registry.h - voice structure ("chain") signature -> render kernels.
//
// Each a2cu_reg_*.cu instantiates a family of kernels and registers them here;
// a2cu_engine.cu looks a bank's structure up at a2cu_bank_new / a2cu_pool_open.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/a2cu.h"
#include "a2cu_kernels.cuh"

typedef void (*render_fn)(const a2cu::RenderParams);
// One instantiation of the warp-specialised kernel (a2cu_split.cuh)
struct SplitVariant {
    render_fn fn;           // nullptr: none
    size_t smem;            // dynamic shared memory without the staged wavetable
    int threads;
    int voices;             // voices per CTA (32 x voice sets)
};
struct KernelEntry {
    render_fn fn;
    int words;      // incl. the flags word
    const char *name;
    // [0]: one voice set per CTA (small banks: most CTAs, shortest critical path);
    // [1]: two voice sets per CTA sharing one staged wavetable (large banks: more voices per SM);
    // [2]: like [0] plus the raw-tap gather for waves without a coefficient table
    SplitVariant split[3];
};
std::map<std::string, KernelEntry> &a2cu_registry();
void a2cu_register_bank_wt();      // render_bank<...>: wavetable chains
void a2cu_register_bank_fm();      // render_bank<...>: FM chains
void a2cu_register_split();        // render_split<...>; needs a current device (function attributes)

static inline std::string sig_of(const a2cu_unitspec *c, int n) {
    std::string s;
    char b[32];
    for (int i = 0; i < n; ++i) {
        snprintf(b, sizeof(b), "%d:%d%d%d%d;", c[i].kind, c[i].ninputs, c[i].noutputs,
                 c[i].add ? 1 : 0, c[i].wireout ? 1 : 0);
        s += b;
    }
    return s;
}
template <class CH>
static void reg_chain(std::vector<a2cu_unitspec> specs, const char *name) {
    KernelEntry e;
    e.fn = a2cu::render_bank<CH>;
    e.words = CH::kWords + 1;
    e.name = name;
    e.split[0] = SplitVariant{nullptr, 0, 0, 0};
    e.split[1] = SplitVariant{nullptr, 0, 0, 0};
    e.split[2] = SplitVariant{nullptr, 0, 0, 0};
    a2cu_registry()[sig_of(specs.data(), (int)specs.size())] = e;
}
// dynamic part; the kernel also has ~4.7 KB static (fused root stage); 227 KB per CTA
static constexpr size_t kMaxSplitSmem = 222 * 1024;

// spec helpers: {kind, nin, nout, add, wireout}
#define S_OSC0 {A2CU_WTOSC, 0, 1, 0, 0}       /* first generator: replaces scratch */
#define S_OSCA {A2CU_WTOSC, 0, 1, 1, 0}       /* further generators: add */
#define S_OSCW {A2CU_WTOSC, 0, 1, 1, 1}       /* lone wtosc, straight to the bus */
#define S_PM12W {A2CU_PANMIX, 1, 2, 1, 1}
#define S_F11 {A2CU_FILTER12, 1, 1, 0, 0}
#define S_F11W {A2CU_FILTER12, 1, 1, 1, 1}
#define S_WS11 {A2CU_WAVESHAPER, 1, 1, 0, 0}
#define S_FM(k) {k, 0, 1, 0, 0}

namespace a2cu {
typedef WtOsc<false, false> Osc0;
typedef WtOsc<true, false> OscA;
typedef WtOsc<true, true> OscW;
typedef PanMix<1, 2, true, true> Pm12W;
typedef Filter12<1, false, false> F11;
typedef Filter12<1, true, true> F11W;
typedef WaveShaper<1, false, false> Ws11;
typedef Fm<1, 0, 0, false, false> Fm1;
typedef Fm<2, 1, 0, false, false> Fm2;
typedef Fm<3, 2, 0, false, false> Fm3;
typedef Fm<4, 2, 0, false, false> Fm4;
typedef Fm<3, 2, 1, false, false> Fm3p;
typedef Fm<4, 2, 1, false, false> Fm4p;
typedef Fm<2, 1, 2, false, false> Fm2r;
typedef Fm<4, 2, 2, false, false> Fm4r;
}  // namespace a2cu
