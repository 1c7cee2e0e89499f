/*
 * amb_private.h - development entry points of libaudiomodem_b200.so that are NOT part of the drop-in boundary
 * (include/audio_modem_b200.h).  They are exported for the test-suite and the profiling scripts only.
 */
#ifndef AMB_PRIVATE_H
#define AMB_PRIVATE_H
#include "../../include/audio_modem_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* development counters: 0 kernel launches, 1 data symbols redone by the float64-faithful fallback of the fast
 * demodulator, 2 link-resolution iterations, 3 link-resolution passes, 4 slotted tiles and 5 candidate tiles of the last push,
 * 6 phase-B links answered by phase A's own walk (its start state was the exact one), 7 phase-B links walked again */
int64_t amb_debug_counter(amb_modem* m, int which);
/* copies the last push's per-slot records to host memory (which = 0: {stream, arena tile, flags, candidate span} int32 x 4;
 * which = 1: {max metric, min Ra, min Rb, excluded} float x 4); returns the slot count */
int64_t amb_debug_read(amb_modem* m, int which, void* out, int64_t cap_bytes);
/* self-tests that need the device's 32-lane execution: which = 0 runs n random batches of the streaming scanner's
 * candidate logic (app.js:806-838) through its warp-parallel form and through the reference-order loop;
 * out2[0] = disagreements (must be 0), out2[1] = batches that committed a preamble */
int amb_selftest(amb_modem* m, int which, uint64_t seed, int n, int64_t* out2);
#ifdef __cplusplus
}
#endif
#endif
