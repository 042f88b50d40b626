// a2cu_reg_bank_fm.cu - render_bank<Chain> instantiations for FM voice structures.
#include "a2cu_registry.h"
using namespace a2cu;

void a2cu_register_bank_fm() {
    reg_chain<Chain<Fm1, Pm12W>>({S_FM(A2CU_FM1), S_PM12W}, "fm1_panmix");
    reg_chain<Chain<Fm2, Pm12W>>({S_FM(A2CU_FM2), S_PM12W}, "fm2_panmix");
    reg_chain<Chain<Fm3, Pm12W>>({S_FM(A2CU_FM3), S_PM12W}, "fm3_panmix");
    reg_chain<Chain<Fm4, Pm12W>>({S_FM(A2CU_FM4), S_PM12W}, "fm4_panmix");
    reg_chain<Chain<Fm3p, Pm12W>>({S_FM(A2CU_FM3P), S_PM12W}, "fm3p_panmix");
    reg_chain<Chain<Fm4p, Pm12W>>({S_FM(A2CU_FM4P), S_PM12W}, "fm4p_panmix");
    reg_chain<Chain<Fm2r, Pm12W>>({S_FM(A2CU_FM2R), S_PM12W}, "fm2r_panmix");
    reg_chain<Chain<Fm4r, Pm12W>>({S_FM(A2CU_FM4R), S_PM12W}, "fm4r_panmix");
    reg_chain<Chain<Fm2, Ws11, Pm12W>>({S_FM(A2CU_FM2), S_WS11, S_PM12W}, "fm2_waveshaper_panmix");
}
