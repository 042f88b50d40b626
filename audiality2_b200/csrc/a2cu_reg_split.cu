// a2cu_reg_split.cu - render_split<...> instantiations (warp-specialised wavetable kernel).
#include "a2cu_registry.h"
#include "a2cu_split.cuh"
using namespace a2cu;

template <int NOSC, bool FILT, int NA>
static void reg_split(std::vector<a2cu_unitspec> specs) {
    KernelEntry &e = a2cu_registry()[sig_of(specs.data(), (int)specs.size())];
    e.split_fn = render_split<NOSC, FILT, NA>;
    e.split_smem = SplitLayout<NOSC, FILT>::bytes;
    e.split_threads = SplitWarps<FILT, NA>::threads;
    cudaFuncSetAttribute(render_split<NOSC, FILT, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)kMaxSplitSmem);
}

void a2cu_register_split() {
    // <oscillators, filter12, helper warps>: the control warp keeps the whole
    // voice in registers, so wider voices get fewer warps per CTA
    reg_split<1, false, 14>({S_OSC0, S_PM12W});
    reg_split<2, false, 14>({S_OSC0, S_OSCA, S_PM12W});
    reg_split<3, false, 10>({S_OSC0, S_OSCA, S_OSCA, S_PM12W});
    reg_split<4, false, 10>({S_OSC0, S_OSCA, S_OSCA, S_OSCA, S_PM12W});
    reg_split<8, false, 6>({S_OSC0, S_OSCA, S_OSCA, S_OSCA, S_OSCA, S_OSCA, S_OSCA, S_OSCA, S_PM12W});
    // with filter12: 11 (8) helpers + control on sub-partitions 0-2, the recurrence alone on 3
    reg_split<1, true, 11>({S_OSC0, S_F11, S_PM12W});
    reg_split<2, true, 11>({S_OSC0, S_OSCA, S_F11, S_PM12W});
    reg_split<3, true, 8>({S_OSC0, S_OSCA, S_OSCA, S_F11, S_PM12W});
}
