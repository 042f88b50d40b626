// a2cu_reg_bank_wt.cu - render_bank<Chain> instantiations for wavetable voice structures.
#include "a2cu_registry.h"
using namespace a2cu;

void a2cu_register_bank_wt() {
    reg_chain<Chain<OscW>>({S_OSCW}, "wtosc");
    reg_chain<Chain<Osc0, Pm12W>>({S_OSC0, S_PM12W}, "wtosc_panmix");
    reg_chain<Chain<Osc0, F11, Pm12W>>({S_OSC0, S_F11, S_PM12W}, "wtosc_filter12_panmix");
    reg_chain<Chain<Osc0, F11W>>({S_OSC0, S_F11W}, "wtosc_filter12");
    reg_chain<Chain<Osc0, OscA, Pm12W>>({S_OSC0, S_OSCA, S_PM12W}, "wtosc2_panmix");
    reg_chain<Chain<Osc0, OscA, OscA, Pm12W>>({S_OSC0, S_OSCA, S_OSCA, S_PM12W}, "wtosc3_panmix");
    reg_chain<Chain<Osc0, OscA, OscA, OscA, Pm12W>>({S_OSC0, S_OSCA, S_OSCA, S_OSCA, S_PM12W}, "wtosc4_panmix");
    reg_chain<Chain<Osc0, OscA, OscA, OscA, OscA, OscA, OscA, OscA, Pm12W>>(
        {S_OSC0, S_OSCA, S_OSCA, S_OSCA, S_OSCA, S_OSCA, S_OSCA, S_OSCA, S_PM12W}, "wtosc8_panmix");
    reg_chain<Chain<Osc0, OscA, F11W>>({S_OSC0, S_OSCA, S_F11W}, "wtosc2_filter12");
    reg_chain<Chain<Osc0, OscA, F11, Pm12W>>({S_OSC0, S_OSCA, S_F11, S_PM12W}, "wtosc2_filter12_panmix");
    reg_chain<Chain<Osc0, OscA, OscA, F11, Pm12W>>({S_OSC0, S_OSCA, S_OSCA, S_F11, S_PM12W},
                                                  "wtosc3_filter12_panmix");
    reg_chain<Chain<Osc0, Ws11, Pm12W>>({S_OSC0, S_WS11, S_PM12W}, "wtosc_waveshaper_panmix");
}
