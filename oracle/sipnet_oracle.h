/*
 * sipnet_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's per-timestep integration loop.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load this.  The product library never links or calls it.
 *
 * Parity is PINNED: tests/test_oracle_vs_reference.py checks this restatement
 * bit-for-bit (every Envi/Fluxes/Trackers field, every step) against the
 * unmodified reference compiled into oracle/_ref (oracle/ref_shim.c) on the
 * reference's four smoke cases and on synthetic event/ensemble cases, and
 * tests/test_oracle_golden.py checks it against the committed golden vectors
 * under tests/golden/ (generated from the reference by
 * tests/golden/make_golden.py).
 */
#ifndef SIPNET_ORACLE_H
#define SIPNET_ORACLE_H

#include <stdint.h>

#include "../include/sipnet_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Run one member (one parameter vector on one site) for the whole climate
 * record.  Same call shape as oracle/ref_shim.c:sipref_run.
 *   flags    : 12 ints, context.h:45-56 order
 *   params   : 80 doubles as left by readParamData() (pre-setupModel)
 *   out32    : [T][32] outputState() columns, or NULL
 *   dbg      : [T][106] debug-log fields, or NULL
 *   recs     : events.out records (capacity max_recs), *nrec = number produced
 * Returns 0 or the reference exit code (3,4,5,7); *steps_done = finished steps.
 */
int sipnet_oracle_run(const int32_t *flags, const double *params, int64_t T,
                      const int32_t *year, const int32_t *day,
                      const double *time, const double *length,
                      const double *tair, const double *tsoil,
                      const double *par, const double *precip,
                      const double *vpd, const double *vpdSoil,
                      const double *vPress, const double *wspd,
                      const double *gdd, int64_t nev,
                      const sipnet_gpu_event *ev, double *out32, double *dbg,
                      int64_t *steps_done, sipnet_gpu_event_record *recs,
                      int32_t max_recs, int32_t *nrec);

/* One member's run returning the mass-balance tracker's check values per step (balance.c:129-148):
 * balance[2 t] = deltaC, balance[2 t + 1] = deltaN.  clim11 = the eleven climate arrays in the order above. */
int sipnet_oracle_run_balance(const int32_t *flags, const double *params, int64_t T,
                              const int32_t *year, const int32_t *day,
                              const double *const *clim11, int64_t nev,
                              const sipnet_gpu_event *ev, double *balance,
                              int64_t *steps_done);
/* ... plus the informational status bits of the run (limitations.c messages). */
int sipnet_oracle_run_diag(const int32_t *flags, const double *params, int64_t T,
                           const int32_t *year, const int32_t *day,
                           const double *const *clim11, int64_t nev,
                           const sipnet_gpu_event *ev, double *balance,
                           int64_t *steps_done, uint32_t *info);

/* occurrence counts (SIPNET_GPU_CNT_* order) of the last _run_diag / _run_balance call on this thread */
void sipnet_oracle_last_counts(uint32_t *out);

/*
 * Ensemble form used by the CPU baseline: run `nmembers` parameter vectors
 * (SoA [80][ld]) on one site with `nthreads` POSIX threads; only the 32
 * output columns of the final step are kept per member ([nmembers][32]) so the
 * timing is of the integration loop, not of a store stream.
 */
int sipnet_oracle_run_ensemble(const int32_t *flags, const double *params_soa,
                               int64_t ld, int64_t nmembers, int64_t T,
                               const int32_t *year, const int32_t *day,
                               const double *const *clim11, int64_t nev,
                               const sipnet_gpu_event *ev, int nthreads,
                               double *final_out32);

#ifdef __cplusplus
}
#endif
#endif
