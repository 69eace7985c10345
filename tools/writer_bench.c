/* writer_bench.c -- rows/s of the sipnet.out row formatter (host/sip_output.c) against the printf statements it
 * reproduces, one thread, model-shaped values.
 *   gcc -O2 -o /tmp/writer_bench tools/writer_bench.c sipnet_b200/host/sip_output.c sipnet_b200/host/sip_config.c \
 *       sipnet_b200/host/sip_inputs.c sipnet_b200/host/sip_restart.c -lm -pthread && /tmp/writer_bench [dir] */
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../sipnet_b200/host/sip_host.h"

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv) {
  enum { ROWS = 200000 };
  static const int kScale[SIPNET_GPU_NOUT] = {4, 2, 0, 4, 3, 2, 3, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -1, -1, 0, 2, 1, 0, -4, -3, -3, -2, -3, 2};
  double *v = (double *)malloc(sizeof(double) * ROWS * SIPNET_GPU_NOUT);
  unsigned long long s = 88172645463325252ull;
  for (size_t i = 0; i < (size_t)ROWS * SIPNET_GPU_NOUT; ++i) {
    s ^= s << 13, s ^= s >> 7, s ^= s << 17;
    double u = (double)(s >> 11) / 9007199254740992.0, sc = 1.0;
    for (int k = 0; k < abs(kScale[i % SIPNET_GPU_NOUT]); ++k) sc *= 10.0;
    v[i] = kScale[i % SIPNET_GPU_NOUT] >= 0 ? u * sc : u / sc;
    if ((s & 7u) == 0) v[i] = 0.0;
  }
  char a[SIP_STATE_ROW_MAX], b[SIP_STATE_ROW_MAX];
  size_t bytes = 0, diff = 0;
  double t0 = now();
  for (int r = 0; r < ROWS; ++r) bytes += sip_format_state_row(a, 2015, 1 + r % 365, 12.0 * (r & 1), v + (size_t)r * SIPNET_GPU_NOUT, 1);
  double t1 = now();
  for (int r = 0; r < ROWS; ++r) bytes += sip_format_state_row_printf(b, 2015, 1 + r % 365, 12.0 * (r & 1), v + (size_t)r * SIPNET_GPU_NOUT, 1);
  double t2 = now();
  for (int r = 0; r < ROWS; ++r) {
    const size_t na = sip_format_state_row(a, 2015, 1 + r % 365, 12.0 * (r & 1), v + (size_t)r * SIPNET_GPU_NOUT, 1);
    const size_t nb = sip_format_state_row_printf(b, 2015, 1 + r % 365, 12.0 * (r & 1), v + (size_t)r * SIPNET_GPU_NOUT, 1);
    if (na != nb || memcmp(a, b, na) != 0) ++diff;
  }
  printf("fast %.3g rows/s (%.0f ns/row), printf %.3g rows/s (%.0f ns/row), speed-up %.1fx, %zu of %d rows differ\n",
         ROWS / (t1 - t0), 1e9 * (t1 - t0) / ROWS, ROWS / (t2 - t1), 1e9 * (t2 - t1) / ROWS, (t2 - t1) / (t1 - t0), diff, ROWS);
  /* with a directory argument: 21 members' files through the thread pool, compared with the printf rows (this is
   * what runs under -fsanitize=thread / address) */
  if (argc > 1) {
    enum { M = 21, T = 300 };
    double *buf = (double *)malloc(sizeof(double) * SIPNET_GPU_NOUT * T * M); /* [col][T][M] */
    for (size_t i = 0; i < (size_t)SIPNET_GPU_NOUT * T * M; ++i) buf[i] = v[i % ((size_t)ROWS * SIPNET_GPU_NOUT)];
    static int32_t year[T], day[T];
    static double tm[T];
    for (int t = 0; t < T; ++t) year[t] = 2011 + t / 100, day[t] = 1 + t % 365, tm[t] = 12.0 * (t & 1);
    char *paths = (char *)calloc(M, SIP_STATE_PATH_MAX);
    int64_t nsteps[M];
    const int32_t *yp[M], *dp[M];
    const double *tp[M];
    for (int m = 0; m < M; ++m) {
      snprintf(paths + (size_t)m * SIP_STATE_PATH_MAX, SIP_STATE_PATH_MAX, "%s/w.out.%d", argv[1], m);
      nsteps[m] = T - 7 * m, yp[m] = year, dp[m] = day, tp[m] = tm;
    }
    const int rc = sip_write_state_files(paths, M, nsteps, yp, dp, tp, T, buf, 1, 4);
    size_t bad = rc != 0;
    for (int m = 0; m < M && !bad; ++m) {
      FILE *f = fopen(paths + (size_t)m * SIP_STATE_PATH_MAX, "r");
      char line[SIP_STATE_ROW_MAX];
      if (!f || !fgets(line, sizeof line, f)) bad++; /* header */
      for (int t = 0; f && t < nsteps[m]; ++t) {
        double row[SIPNET_GPU_NOUT];
        for (int c = 0; c < SIPNET_GPU_NOUT; ++c) row[c] = buf[((size_t)c * T + (size_t)t) * M + (size_t)m];
        const size_t nb = sip_format_state_row_printf(b, year[t], day[t], tm[t], row, 1);
        b[nb] = 0;
        if (!fgets(line, sizeof line, f) || strcmp(line, b) != 0) bad++;
      }
      if (f && fgets(line, sizeof line, f)) bad++; /* nothing after the last row */
      if (f) fclose(f);
    }
    printf("thread pool: rc %d, %zu mismatches over %d members\n", rc, bad, M);
    free(buf);
    free(paths);
    diff += bad;
  }
  free(v);
  return diff != 0 || bytes == 0;
}
