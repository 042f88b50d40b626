// a2cu_reg_split.cu - render_split<...> instantiations (warp-specialised wavetable kernel).
#include "a2cu_registry.h"
#include "a2cu_split.cuh"
using namespace a2cu;

// room the staged Hermite-coefficient table of a builtin 2048-point wave needs (all mip levels)
static constexpr size_t kTableRoom = 90 * 1024;

template <int NOSC, bool FILT, int NH, int VS, int R, bool RAW>
static void reg_variant(KernelEntry &e, int slot) {
    e.split[slot].fn = render_split<NOSC, FILT, NH, VS, R, RAW>;
    e.split[slot].smem = split_smem_bytes<NOSC, FILT, R, VS>();
    e.split[slot].threads = SplitWarps<FILT, NH, VS>::threads;
    e.split[slot].voices = 32 * VS;
    cudaFuncSetAttribute(render_split<NOSC, FILT, NH, VS, R, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)kMaxSplitSmem);
}

// <oscillators, filter12, helper warps for one set per CTA, helper warps per set for two sets per CTA>
template <int NOSC, bool FILT, int NH1, int NH2>
static void reg_split(std::vector<a2cu_unitspec> specs) {
    KernelEntry &e = a2cu_registry()[sig_of(specs.data(), (int)specs.size())];
    reg_variant<NOSC, FILT, NH1, 1, 4, false>(e, 0);
    reg_variant<NOSC, FILT, NH1, 1, 4, true>(e, 2);     // banks that play table-less (large sampled) waves
    // two voice sets: pays where the recurrence warp is the critical path (helpers have slack) and
    // both sets and the table fit the 227 KB of one CTA
    if constexpr (!FILT) return;
    else if constexpr (split_smem_bytes<NOSC, FILT, 4, 2>() + kTableRoom <= kMaxSplitSmem)
        reg_variant<NOSC, FILT, NH2, 2, 4, false>(e, 1);
    else if constexpr (split_smem_bytes<NOSC, FILT, 3, 2>() + kTableRoom <= kMaxSplitSmem)
        reg_variant<NOSC, FILT, NH2, 2, 3, false>(e, 1);
}

void a2cu_register_split() {
    // the control warp keeps the whole voice in registers, so wider voices get fewer warps per CTA
    reg_split<1, false, 14, 7>({S_OSC0, S_PM12W});
    reg_split<2, false, 14, 7>({S_OSC0, S_OSCA, S_PM12W});
    reg_split<3, false, 10, 5>({S_OSC0, S_OSCA, S_OSCA, S_PM12W});
    reg_split<4, false, 10, 5>({S_OSC0, S_OSCA, S_OSCA, S_OSCA, S_PM12W});
    reg_split<8, false, 6, 3>({S_OSC0, S_OSCA, S_OSCA, S_OSCA, S_OSCA, S_OSCA, S_OSCA, S_OSCA, S_PM12W});
    // with filter12: helpers + control on sub-partitions 0-2, the recurrence warps alone on 3
    reg_split<1, true, 11, 5>({S_OSC0, S_F11, S_PM12W});
    reg_split<2, true, 11, 5>({S_OSC0, S_OSCA, S_F11, S_PM12W});
    reg_split<3, true, 8, 5>({S_OSC0, S_OSCA, S_OSCA, S_F11, S_PM12W});
}
