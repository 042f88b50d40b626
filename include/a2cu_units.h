/*
 * a2cu_units.h - the drop-in boundary: what liba2cu_units.so exports.
 *
 * These are exactly the symbols the reference binds for its hot-path units:
 * a2_core_units[] (src/audiality2.c:183-207) takes the address of each
 * descriptor and hands it to a2_RegisterUnit() (src/units.c:79-157). The
 * declarations below replace, one for one:
 *
 *   src/units/wtosc.h:28       a2_wtosc_unitdesc
 *   src/units/panmix.h:28      a2_panmix_unitdesc
 *   src/units/filter12.h:28    a2_filter12_unitdesc
 *   src/units/waveshaper.h:28  a2_waveshaper_unitdesc
 *   src/units/fm.h:28-37       a2_fm1 / fm2 / fm3 / fm4 / fm3p / fm4p / fm2r /
 *                              fm4r _unitdesc
 *   src/units/inline.h:40      a2_inline_unitdesc (the compiler compares its
 *                              address, src/compiler.c:3029; our version
 *                              brackets each sub-tree's bus on the device and
 *                              calls the host's a2_inline_ProcessAdd,
 *                              src/core.c:1763-1767, for the recursion)
 *   src/units/fbdelay.h:28     a2_fbdelay_unitdesc (the
 *                              effect every song chains right after its
 *                              mix-down; on the host it would force a device
 *                              round trip per song and fragment)
 *   src/units/limiter.h:28     a2_limiter_unitdesc   } the other bus effects of
 *   src/units/dcblock.h:28     a2_dcblock_unitdesc   } SURVEY.md 8(f)1, same
 *   src/units/dc.h:28          a2_dc_unitdesc        } reason
 *
 * Each descriptor carries the reference's name, flags, register names in VM
 * register order, constants and I/O limits (include/a2_units.h:225-252), so
 * A2S scripts compile unchanged. Callback contracts: include/a2_units.h:115
 * (write), :132 (Initialize), :142 (Deinitialize), :159-160 (Open/CloseState),
 * :176 (Process).
 *
 * Build the host as the reference minus src/units/{wtosc,panmix,filter12,fm,
 * waveshaper,inline,fbdelay,limiter,dcblock,dc}.c and link this library in their place.
 *
 * a2cu_RegisterDriver() additionally registers a "cuda" audio driver
 * (a2_RegisterDriver, include/a2_drivers.h:193) that behaves like the
 * reference's "buffer" driver (src/drivers/bufferdrv.c:28-110); select it with
 * a2_NewDriver(A2_AUDIODRIVER, "cuda") / a2play -dcuda.
 *
 * Environment: A2CU_DEVICE = CUDA ordinal (default 0).
 */
#ifndef A2CU_UNITS_H
#define A2CU_UNITS_H

#ifdef __cplusplus
extern "C" {
#endif

struct A2_unitdesc;

extern const struct A2_unitdesc a2_wtosc_unitdesc;
extern const struct A2_unitdesc a2_panmix_unitdesc;
extern const struct A2_unitdesc a2_filter12_unitdesc;
extern const struct A2_unitdesc a2_waveshaper_unitdesc;
extern const struct A2_unitdesc a2_fm1_unitdesc;
extern const struct A2_unitdesc a2_fm2_unitdesc;
extern const struct A2_unitdesc a2_fm3_unitdesc;
extern const struct A2_unitdesc a2_fm4_unitdesc;
extern const struct A2_unitdesc a2_fm3p_unitdesc;
extern const struct A2_unitdesc a2_fm4p_unitdesc;
extern const struct A2_unitdesc a2_fm2r_unitdesc;
extern const struct A2_unitdesc a2_fm4r_unitdesc;
extern const struct A2_unitdesc a2_inline_unitdesc;
extern const struct A2_unitdesc a2_fbdelay_unitdesc;
extern const struct A2_unitdesc a2_limiter_unitdesc;
extern const struct A2_unitdesc a2_dcblock_unitdesc;
extern const struct A2_unitdesc a2_dc_unitdesc;

/* Returns an A2_errors code (0 = A2_OK). */
int a2cu_RegisterDriver(void);

#ifdef __cplusplus
}
#endif
#endif /* A2CU_UNITS_H */
