/*
 * sip_inputs.c -- <prefix>.param, <prefix>.clim and events.in readers.
 *
 * Behavioural restatement of readParamData()/readModelParams() (reference
 * src/sipnet/sipnet.c:290-427, src/common/modelParams.c:136-230), readClimData()
 * (sipnet.c:128-277) and readEventData()/createEventNode() (src/sipnet/events.c:39-367):
 * same accepted formats, same unit conversions and floors, same failure codes.
 * Results are flat arrays (the device wants arrays, not linked lists).
 */
#define _GNU_SOURCE
#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <unistd.h>

#include "sip_host.h"

int sip_fail(int code, const char *fmt, ...);
void sip_info(int quiet, const char *fmt, ...);

#define SIP_TINY 0.000001 /* util.h:14 */

/* ---- parameters ----------------------------------------------------------------------------- */
typedef struct {
  const char *name; /* as spelled in the .param file */
  int index;        /* row in struct Parameters order */
  int required;
} param_spec;

int sip_read_params(const char *path, const sipnet_gpu_flags *f, int quiet, double out[SIPNET_GPU_NPARAMS]) {
  const int litter = f->litterPool, ncyc = f->nitrogenCycle, anaer = f->anaerobic;
  /* registration order and "required" rules of readParamData(), sipnet.c:300-396 */
  const param_spec specs[] = {
      {"plantWoodInit", SIPNET_P_plantWoodInit, 1},
      {"laiInit", SIPNET_P_laiInit, 1},
      {"litterInit", SIPNET_P_litterInit, 1},
      {"soilInit", SIPNET_P_soilInit, 1},
      {"soilWFracInit", SIPNET_P_soilWFracInit, 1},
      {"snowInit", SIPNET_P_snowInit, 1},
      {"aMax", SIPNET_P_aMax, 1},
      {"aMaxFrac", SIPNET_P_aMaxFrac, 1},
      {"baseFolRespFrac", SIPNET_P_baseFolRespFrac, 1},
      {"psnTMin", SIPNET_P_psnTMin, 1},
      {"psnTOpt", SIPNET_P_psnTOpt, 1},
      {"vegRespQ10", SIPNET_P_vegRespQ10, 1},
      {"growthRespFrac", SIPNET_P_growthRespFrac, f->growthResp},
      {"frozenSoilFolREff", SIPNET_P_frozenSoilFolREff, 1},
      {"frozenSoilThreshold", SIPNET_P_frozenSoilThreshold, 1},
      {"dVpdSlope", SIPNET_P_dVpdSlope, 1},
      {"dVpdExp", SIPNET_P_dVpdExp, 1},
      {"halfSatPar", SIPNET_P_halfSatPar, 1},
      {"attenuation", SIPNET_P_attenuation, 1},
      {"leafOnDay", SIPNET_P_leafOnDay, !(f->gdd || f->soilPhenol)},
      {"gddLeafOn", SIPNET_P_gddLeafOn, f->gdd},
      {"soilTempLeafOn", SIPNET_P_soilTempLeafOn, f->soilPhenol},
      {"leafOffDay", SIPNET_P_leafOffDay, 1},
      {"leafGrowth", SIPNET_P_leafGrowth, 1},
      {"fracLeafFall", SIPNET_P_fracLeafFall, 1},
      {"leafAllocation", SIPNET_P_leafAllocation, 1},
      {"leafTurnoverRate", SIPNET_P_leafTurnoverRate, 1},
      {"baseVegResp", SIPNET_P_baseVegResp, 1},
      {"litterBreakdownRate", SIPNET_P_litterBreakdownRate, litter},
      {"fracLitterRespired", SIPNET_P_fracLitterRespired, litter},
      {"baseSoilResp", SIPNET_P_baseSoilResp, 1},
      {"soilRespQ10", SIPNET_P_soilRespQ10, 1},
      {"soilRespMoistEffect", SIPNET_P_soilRespMoistEffect, f->waterHResp},
      {"waterRemoveFrac", SIPNET_P_waterRemoveFrac, 1},
      {"frozenSoilEff", SIPNET_P_frozenSoilEff, 1},
      {"wueConst", SIPNET_P_wueConst, 1},
      {"soilWHC", SIPNET_P_soilWHC, 1},
      {"immedEvapFrac", SIPNET_P_immedEvapFrac, 1},
      {"fastFlowFrac", SIPNET_P_fastFlowFrac, 1},
      {"leafPoolDepth", SIPNET_P_leafPoolDepth, f->leafWater},
      {"snowMelt", SIPNET_P_snowMelt, f->snow},
      {"rdConst", SIPNET_P_rdConst, 1},
      {"rSoilConst1", SIPNET_P_rSoilConst1, 1},
      {"rSoilConst2", SIPNET_P_rSoilConst2, 1},
      {"leafCSpWt", SIPNET_P_leafCSpWt, 1},
      {"cFracLeaf", SIPNET_P_cFracLeaf, 1},
      {"woodTurnoverRate", SIPNET_P_woodTurnoverRate, 1},
      {"fineRootFrac", SIPNET_P_fineRootFrac, 1},
      {"coarseRootFrac", SIPNET_P_coarseRootFrac, 1},
      {"fineRootAllocation", SIPNET_P_fineRootAllocation, 1},
      {"woodAllocation", SIPNET_P_woodAllocation, 1},
      {"fineRootTurnoverRate", SIPNET_P_fineRootTurnoverRate, 1},
      {"coarseRootTurnoverRate", SIPNET_P_coarseRootTurnoverRate, 1},
      {"baseFineRootResp", SIPNET_P_baseFineRootResp, 1},
      {"baseCoarseRootResp", SIPNET_P_baseCoarseRootResp, 1},
      {"fineRootQ10", SIPNET_P_fineRootQ10, 1},
      {"coarseRootQ10", SIPNET_P_coarseRootQ10, 1},
      {"mineralNInit", SIPNET_P_minNInit, ncyc},
      {"soilOrgNInit", SIPNET_P_soilOrgNInit, ncyc},
      {"litterOrgNInit", SIPNET_P_litterOrgNInit, ncyc},
      {"plantStorageNInit", SIPNET_P_plantStorageNInit, ncyc},
      {"nVolatilizationFrac", SIPNET_P_nVolatilizationFrac, ncyc},
      {"nLeachingFrac", SIPNET_P_nLeachingFrac, ncyc},
      {"leafCN", SIPNET_P_leafCN, ncyc},
      {"woodCN", SIPNET_P_woodCN, ncyc},
      {"fineRootCN", SIPNET_P_fineRootCN, ncyc},
      {"kCN", SIPNET_P_kCN, ncyc},
      {"nFixationFracMax", SIPNET_P_nFixationFracMax, ncyc},
      {"halfNFixationMax", SIPNET_P_halfNFixationMax, ncyc},
      {"leafOnReallocFrac", SIPNET_P_leafOnReallocFrac, 1},
      {"leafNResorptionFrac", SIPNET_P_leafNResorptionFrac, ncyc},
      {"fAnoxia", SIPNET_P_fAnoxia, anaer || ncyc},
      {"anaerobicDecompRate", SIPNET_P_anaerobicDecompRate, anaer},
      {"anaerobicTransExp", SIPNET_P_anaerobicTransExp, anaer},
      {"soilMethaneRate", SIPNET_P_soilMethaneRate, anaer},
      {"litterMethaneRate", SIPNET_P_litterMethaneRate, anaer},
      {"waterDrainFrac", SIPNET_P_waterDrainFrac, f->flooding},
      {"soilCSaturation", SIPNET_P_soilCSaturation, f->carbonSaturation},
  };
  const int nspec = (int)(sizeof specs / sizeof specs[0]);
  char seen[sizeof specs / sizeof specs[0]];
  memset(seen, 0, sizeof seen);
  for (int k = 0; k < SIPNET_GPU_NPARAMS; ++k) out[k] = 0.0; /* unread optional parameters stay 0 (zeroed global) */

  FILE *in = fopen(path, "r");
  if (!in) {
    fprintf(stderr, "Error reading '%s': %s\n", path, strerror(errno));
    return sip_fail(SIPNET_GPU_ERR_FILE_OPEN, "cannot open %s", path);
  }
  char line[256], unknown[2048] = "";
  int formatChecked = 0, rc = 0;
  while (rc == 0 && fgets(line, sizeof line, in) != NULL) {
    char *bang = strpbrk(line, "!");
    if (bang) *bang = '\0';
    if (strlen(line) == strspn(line, " \t\n\r")) continue;
    if (!formatChecked) { /* checkParamFormat(), modelParams.c:127-134 */
      char copy[256];
      strcpy(copy, line);
      int nf = 0;
      char *sv = NULL; /* strtok_r: sites are read on several threads in many-site launches */
      for (char *t = strtok_r(copy, " \t\n\r", &sv); t; t = strtok_r(NULL, " \t\n\r", &sv)) ++nf;
      if (nf > 2) sip_info(quiet, "extra columns in .param file are being ignored (found %d columns)\n", nf);
      formatChecked = 1;
    }
    char *sv = NULL;
    char *name = strtok_r(line, " \t\n\r", &sv);
    char *val = strtok_r(NULL, " \t\n\r", &sv);
    if (!name || !val) {
      rc = sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "reading parameter file: missing value for %s", name ? name : "?");
      break;
    }
    if (strcmp(val, "*") == 0) {
      rc = sip_fail(SIPNET_GPU_ERR_BAD_PARAMETER_VALUE, "reading parameter %s; '*' is no longer supported", name);
      break;
    }
    const double v = strtod(val, NULL);
    int hit = -1;
    for (int k = 0; k < nspec && hit < 0; ++k)
      if (strcasecmp(name, specs[k].name) == 0) hit = k;
    if (hit < 0) {
      if (strlen(unknown) + strlen(name) + 3 > sizeof unknown) {
        rc = sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "Too many unknown params; please remove some from %s and rerun", path);
        break;
      }
      if (unknown[0]) strcat(unknown, ", ");
      strcat(unknown, name);
    } else if (seen[hit]) {
      rc = sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "reading parameter file: read %s, but this parameter has already been set",
                    name);
    } else {
      out[specs[hit].index] = v;
      seen[hit] = 1;
    }
  }
  fclose(in);
  if (rc) return rc;
  if (unknown[0]) sip_info(quiet, "Unknown param(s) found (and ignored): %s\n", unknown);
  /* checkAllRead(), modelParams.c:36-70 */
  char missing[1024] = "";
  int missingOpt = 0;
  for (int k = 0; k < nspec; ++k) {
    if (seen[k]) continue;
    if (specs[k].required) {
      if (strlen(missing) + strlen(specs[k].name) + 2 < sizeof missing) {
        strcat(missing, " ");
        strcat(missing, specs[k].name);
      }
    } else {
      missingOpt = 1;
    }
  }
  if (missingOpt && !quiet) {
    fputs("[INFO   ] optional params not specified in input file:", stdout);
    for (int k = 0; k < nspec; ++k)
      if (!seen[k] && !specs[k].required) printf(" %s", specs[k].name);
    fputs("\n", stdout);
  }
  if (missing[0]) return sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "Did not find required parameter(s):%s", missing);
  /* divisor floors, sipnet.c:404-424 */
  const int floors[] = {SIPNET_P_cFracLeaf, SIPNET_P_halfSatPar, SIPNET_P_soilWHC, SIPNET_P_leafCSpWt,
                        SIPNET_P_leafCN,    SIPNET_P_woodCN,     SIPNET_P_fineRootCN};
  for (size_t k = 0; k < sizeof floors / sizeof floors[0]; ++k)
    if (out[floors[k]] < SIP_TINY) out[floors[k]] = SIP_TINY;
  return 0;
}

/* ---- climate ------------------------------------------------------------------------------------ */
static int site_reserve(sip_site_data *s, int64_t cap) {
  int32_t **ip[] = {&s->year, &s->day};
  double **dp[] = {&s->time, &s->length, &s->tair, &s->tsoil, &s->par, &s->precip,
                   &s->vpd,  &s->vpdSoil, &s->vPress, &s->wspd, &s->gdd};
  for (size_t k = 0; k < 2; ++k) {
    int32_t *n = (int32_t *)realloc(*ip[k], (size_t)cap * sizeof(int32_t));
    if (!n) return -1;
    *ip[k] = n;
  }
  for (size_t k = 0; k < 11; ++k) {
    double *n = (double *)realloc(*dp[k], (size_t)cap * sizeof(double));
    if (!n) return -1;
    *dp[k] = n;
  }
  return 0;
}

/* ---- fast path of the .clim reader ---------------------------------------------------------------------------
 * readClimData() reads the file with fscanf("%d %d %lf ... %lf") -- a token stream in which line ends are just
 * white space.  Ten-year half-daily sites are 7 306 records of 12 numbers, a 10 000-site launch 73 million records,
 * and glibc's scanf needs ~150 ns per number.  A record that sits on ONE line as exactly 12 plainly written numbers
 * (the only way the files are ever written) is converted directly; at the first line that is anything else -- fewer
 * or more tokens, a token scanf would split or reject, "nan", hex floats, an over-long line -- the file is
 * repositioned to the start of that line and the fscanf loop carries on from there, so every irregular file is read
 * exactly as before.
 * Numbers: "%d" of [+-]digits (at most 9) is that integer; "%lf" of [+-]digits[.digits][e[+-]digits] is strtod's
 * correctly rounded value, which for at most 15 significant digits and a decimal exponent within +-22 is ONE
 * correctly rounded multiplication or division of two exactly representable numbers (Clinger's fast path);
 * everything else goes through strtod itself. */
static const double kExact10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                    1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

static int is_blank_char(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

/* [+-]digits{1,9} followed by white space; returns the end of the token or NULL */
static const char *plain_int(const char *p, int *out) {
  int neg = 0, n = 0;
  long v = 0;
  if (*p == '+' || *p == '-') neg = (*p++ == '-');
  while (*p >= '0' && *p <= '9' && n < 10) v = v * 10 + (*p++ - '0'), ++n;
  if (n == 0 || n > 9 || !is_blank_char(*p)) return NULL;
  *out = (int)(neg ? -v : v);
  return p;
}

/* [+-](digits[.digits*] | .digits)[(e|E)[+-]digits] followed by white space; returns the end of the token or NULL */
static const char *plain_double(const char *p, double *out) {
  const char *start = p;
  int neg = 0;
  if (*p == '+' || *p == '-') neg = (*p++ == '-');
  uint64_t mant = 0;
  int ndig = 0, nsig = 0, dec = 0; /* digits seen, significant digits kept in mant, digits after the point */
  while (*p >= '0' && *p <= '9') {
    if (nsig < 19 && (mant != 0 || *p != '0')) mant = mant * 10u + (uint64_t)(*p - '0'), ++nsig;
    else if (mant != 0 || *p != '0') nsig = 100; /* more digits than the shortcut handles */
    ++p, ++ndig;
  }
  if (*p == '.') {
    ++p;
    while (*p >= '0' && *p <= '9') {
      if (nsig < 19 && (mant != 0 || *p != '0')) mant = mant * 10u + (uint64_t)(*p - '0'), ++nsig;
      else if (mant != 0 || *p != '0') nsig = 100;
      ++p, ++ndig, ++dec;
    }
  }
  if (ndig == 0) return NULL;
  int e10 = 0;
  if (*p == 'e' || *p == 'E') {
    const char *q = p + 1;
    int eneg = 0, en = 0, ev = 0;
    if (*q == '+' || *q == '-') eneg = (*q++ == '-');
    while (*q >= '0' && *q <= '9' && en < 5) ev = ev * 10 + (*q++ - '0'), ++en;
    if (en == 0 || en > 4 || (*q >= '0' && *q <= '9')) return NULL; /* "1e", "1e+": scanf's business */
    e10 = eneg ? -ev : ev;
    p = q;
  }
  if (!is_blank_char(*p)) return NULL;
  e10 -= dec;
  if (nsig <= 15 && e10 >= -22 && e10 <= 22) { /* mant < 10^15 < 2^53 and 10^|e10| are exact */
    const double m = (double)mant;
    const double v = e10 < 0 ? m / kExact10[-e10] : m * kExact10[e10];
    *out = neg ? -v : v;
    return p;
  }
  char *endp = NULL;
  const double v = strtod(start, &endp);
  if (endp != p) return NULL;
  *out = v;
  return p;
}

/* one record on one line: 12 plain numbers and nothing else.  1 = converted, 0 = not a plain record line */
static int plain_clim_line(const char *p, int *year, int *day, double v[10]) {
  while (*p == ' ' || *p == '\t') ++p;
  if (!(p = plain_int(p, year))) return 0;
  while (*p == ' ' || *p == '\t') ++p;
  if (!(p = plain_int(p, day))) return 0;
  for (int k = 0; k < 10; ++k) {
    while (*p == ' ' || *p == '\t') ++p;
    if (!(p = plain_double(p, &v[k]))) return 0;
  }
  while (*p == ' ' || *p == '\t' || *p == '\r') ++p;
  return *p == '\n' || *p == '\0';
}

int sip_read_clim(const char *path, int gddFlag, int quiet, sip_site_data *s) {
  memset(s, 0, sizeof *s);
  FILE *in = fopen(path, "r");
  if (!in) {
    fprintf(stderr, "Error reading '%s': %s\n", path, strerror(errno));
    return sip_fail(SIPNET_GPU_ERR_FILE_OPEN, "cannot open %s", path);
  }
  char *first = NULL;
  size_t cap0 = 0;
  if (getline(&first, &cap0, in) == -1) {
    free(first);
    fclose(in);
    return sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "no climate data in %s", path);
  }
  /* 12 columns, or the legacy 14 (location first, soilWetness last), sipnet.c:157-178 */
  int nf = 0;
  {
    char *copy = strdup(first), *sv = NULL;
    for (char *t = strtok_r(copy, " \t\n\r", &sv); t; t = strtok_r(NULL, " \t\n\r", &sv)) ++nf;
    free(copy);
  }
  int legacy;
  if (nf == 12) {
    legacy = 0;
  } else if (nf == 14) {
    legacy = 1;
    sip_info(quiet, "old climate file format detected (found %d cols); ignoring location and soilWetness columns in %s\n",
             nf, path);
  } else {
    free(first);
    fclose(in);
    return sip_fail(SIPNET_GPU_ERR_INPUT_FILE,
                    "format unrecognized in climate file %s; %d columns found, expected 12 or 14 (legacy format)", path, nf);
  }
  const int expected = legacy ? 14 : 12;
  int year, day, loc0 = 0, loc = 0, status;
  double time, length, tair, tsoil, par, precip, vpd, vpdSoil, vPress, wspd, wet;
  if (legacy)
    status = sscanf(first, "%d %d %d %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf", &loc0, &year, &day, &time, &length,
                    &tair, &tsoil, &par, &precip, &vpd, &vpdSoil, &vPress, &wspd, &wet);
  else
    status = sscanf(first, "%d %d %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf", &year, &day, &time, &length, &tair, &tsoil,
                    &par, &precip, &vpd, &vpdSoil, &vPress, &wspd);
  free(first);
  if (status != expected) {
    fclose(in);
    return sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "while reading climate file: bad data on first line");
  }
  int64_t cap = 0, n = 0;
  int rc = 0;
  int plain = !legacy && getenv("SIPNET_HOST_SCANF_CLIM") == NULL; /* (the variable forces the fscanf loop: tests) */
  long lineStart = plain ? ftell(in) : -1; /* just behind the first line */
  while (status != EOF) {
    if (n == cap) {
      cap = cap ? cap * 2 : 8192;
      if (site_reserve(s, cap)) {
        rc = sip_fail(SIPNET_GPU_ERR_INTERNAL, "memory allocation failure in file processing");
        break;
      }
    }
    /* unit conversions and floors, sipnet.c:205-238 */
    if (length < 0) length = length / -86400.;
    s->year[n] = year;
    s->day[n] = day;
    s->time[n] = time;
    s->length[n] = length;
    s->tair[n] = tair;
    s->tsoil[n] = tsoil;
    s->par[n] = par * (1.0 / length);
    s->precip[n] = precip * 0.1;
    s->vpd[n] = vpd * 0.001;
    if (s->vpd[n] < SIP_TINY) s->vpd[n] = SIP_TINY;
    s->vpdSoil[n] = vpdSoil * 0.001;
    s->vPress[n] = vPress * 0.001;
    s->wspd[n] = wspd;
    if (s->wspd[n] < SIP_TINY) s->wspd[n] = SIP_TINY;
    if (gddFlag) {
      double g = tair * length;
      if (g < 0) g = 0;
      s->gdd[n] = g;
    } else {
      s->gdd[n] = 0.0;
    }
    ++n;
    if (plain) { /* the next record, if it is one plain line (see plain_clim_line) */
      char line[1024];
      const long at = lineStart; /* file offset of this line: the first line's end plus the lines converted since */
      double v[10];
      int got = 0;
      if (at >= 0 && fgets(line, sizeof line, in) != NULL) {
        const size_t len = strlen(line); /* (a line with an embedded NUL looks cut short and is left to fscanf) */
        if (len > 0 && line[len - 1] == '\n' && plain_clim_line(line, &year, &day, v)) {
          time = v[0], length = v[1], tair = v[2], tsoil = v[3], par = v[4], precip = v[5], vpd = v[6], vpdSoil = v[7],
          vPress = v[8], wspd = v[9];
          lineStart += (long)len;
          got = 1;
        }
      }
      if (got) continue; /* status is still `expected` */
      plain = 0;         /* anything else, the end of the file included: fscanf from the start of that line on */
      if (at < 0 || fseek(in, at, SEEK_SET) != 0) {
        rc = sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "while reading climate file: cannot reposition %s", path);
        break;
      }
    }
    if (legacy)
      status = fscanf(in, "%d %d %d %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf", &loc, &year, &day, &time, &length, &tair,
                      &tsoil, &par, &precip, &vpd, &vpdSoil, &vPress, &wspd, &wet);
    else
      status = fscanf(in, "%d %d %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf", &year, &day, &time, &length, &tair, &tsoil,
                      &par, &precip, &vpd, &vpdSoil, &vPress, &wspd);
    if (status != EOF) {
      if (status != expected) {
        rc = sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "while reading climate file: bad data near year %d day %d", year, day);
        break;
      }
      if (legacy && loc != loc0) {
        rc = sip_fail(SIPNET_GPU_ERR_INPUT_FILE,
                      "while reading legacy climate file %s: multiple locations not supported (locations found: %d and %d)",
                      path, loc0, loc);
        break;
      }
    }
  }
  fclose(in);
  s->nsteps = n;
  if (rc) sip_site_free(s);
  return rc;
}

/* ---- events ----------------------------------------------------------------------------------------- */
static int event_type_of(const char *name) { /* eventStringToType(), events.c:210-236 */
  static const struct {
    const char *n;
    int t;
  } map[] = {{"irrig", SIPNET_EV_IRRIGATION}, {"fert", SIPNET_EV_FERTILIZATION}, {"plant", SIPNET_EV_PLANTING},
             {"till", SIPNET_EV_TILLAGE},     {"harv", SIPNET_EV_HARVEST},       {"leafon", SIPNET_EV_LEAFON},
             {"leafoff", SIPNET_EV_LEAFOFF},  {"plantdeath", SIPNET_EV_PLANTDEATH}};
  for (size_t k = 0; k < sizeof map / sizeof map[0]; ++k)
    if (strcmp(name, map[k].n) == 0) return map[k].t;
  return -1;
}

const char *sip_event_type_name(int type) { /* eventTypeToString(), events.c:186-208 */
  switch (type) {
    case SIPNET_EV_IRRIGATION: return "irrig";
    case SIPNET_EV_PLANTING: return "plant";
    case SIPNET_EV_HARVEST: return "harv";
    case SIPNET_EV_FERTILIZATION: return "fert";
    case SIPNET_EV_TILLAGE: return "till";
    case SIPNET_EV_LEAFON: return "leafon";
    case SIPNET_EV_LEAFOFF: return "leafoff";
    case SIPNET_EV_PLANTDEATH: return "plantdeath";
    default: return NULL;
  }
}

/* createEventNode(), events.c:39-184: parse the per-type parameter text */
static int parse_event_params(int year, int day, int type, const char *txt, sipnet_gpu_event *ev) {
  memset(ev, 0, sizeof *ev);
  ev->year = year;
  ev->day = day;
  ev->type = type;
  double a, b, c, d;
  int m;
  switch (type) {
    case SIPNET_EV_HARVEST:
      if (sscanf(txt, "%lf %lf %lf %lf", &a, &b, &c, &d) != 4)
        return sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "parsing Harvest params for year %d day %d", year, day);
      if ((a + c > 1) || (b + d > 1))
        return sip_fail(SIPNET_GPU_ERR_BAD_PARAMETER_VALUE,
                        "invalid harvest newEvent for year %d day %d; above and below must each add to 1 or less", year, day);
      ev->p[0] = a, ev->p[1] = b, ev->p[2] = c, ev->p[3] = d; /* removedAbove, removedBelow, transferredAbove, transferredBelow */
      return 0;
    case SIPNET_EV_IRRIGATION:
      if (sscanf(txt, "%lf %d", &a, &m) != 2)
        return sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "parsing Irrigation params for year %d day %d", year, day);
      ev->p[0] = a;
      ev->method = m;
      return 0;
    case SIPNET_EV_FERTILIZATION:
      if (sscanf(txt, "%lf %lf %lf", &a, &b, &c) != 3)
        return sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "parsing Fertilization params for year %d day %d", year, day);
      ev->p[0] = a, ev->p[1] = b, ev->p[2] = c; /* orgN, orgC, minN */
      return 0;
    case SIPNET_EV_PLANTING:
      if (sscanf(txt, "%lf %lf %lf %lf", &a, &b, &c, &d) != 4)
        return sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "parsing Planting params for year %d day %d", year, day);
      ev->p[0] = a, ev->p[1] = b, ev->p[2] = c, ev->p[3] = d; /* leafC, woodC, fineRootC, coarseRootC */
      return 0;
    case SIPNET_EV_TILLAGE:
      if (sscanf(txt, "%lf", &a) != 1)
        return sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "parsing Tillage params for year %d day %d", year, day);
      ev->p[0] = a;
      return 0;
    case SIPNET_EV_LEAFON:
    case SIPNET_EV_LEAFOFF:
      if (sscanf(txt, "%lf", &a) > 0) /* takes no parameters: any number is an error */
        return sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "parsing %s params for year %d day %d",
                        type == SIPNET_EV_LEAFON ? "LeafOn" : "LeafOff", year, day);
      return 0;
    case SIPNET_EV_PLANTDEATH:
      return sip_fail(SIPNET_GPU_ERR_INPUT_FILE,
                      "PLANTDEATH event found for year %d day %d, but not implemented as an input event", year, day);
    default:
      return sip_fail(SIPNET_GPU_ERR_UNKNOWN_EVENT, "found unknown event type %d while reading event file", type);
  }
}

int sip_read_events(const char *path, const sipnet_gpu_flags *f, const double params[SIPNET_GPU_NPARAMS], int quiet,
                    sip_site_data *s) {
  s->nevents = 0;
  free(s->events);
  s->events = NULL;
  if (access(path, F_OK) != 0) { /* no file is fine, events.c:275-280 */
    sip_info(quiet, "No event file found, assuming no input events\n");
    return 0;
  }
  sip_info(quiet, "Begin reading event data from file %s\n", path);
  FILE *in = fopen(path, "r");
  if (!in) {
    fprintf(stderr, "Error reading '%s': %s\n", path, strerror(errno));
    return sip_fail(SIPNET_GPU_ERR_FILE_OPEN, "cannot open %s", path);
  }
  char line[1024];
  int64_t cap = 0;
  int lastYear = 0, lastDay = 0, leafChecked = 0, rc = 0;
  while (rc == 0 && fgets(line, sizeof line, in) != NULL) {
    const size_t len = strlen(line);
    if (len == sizeof line - 1 && line[len - 1] != '\n') { /* checkEventLineTruncation(), events.c:243-249 */
      rc = sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "Event line too long (exceeds %d chars), data may be truncated", 1024);
      break;
    }
    int year, day, used = 0;
    char typeStr[32];
    if (sscanf(line, "%d %d %31s %n", &year, &day, typeStr, &used) != 3) {
      rc = s->nevents == 0
               ? sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "reading event file: bad data on first line")
               : sip_fail(SIPNET_GPU_ERR_INPUT_FILE, "reading event file: bad data on line after year %d day %d", lastYear,
                          lastDay);
      break;
    }
    const int type = event_type_of(typeStr);
    if (type < 0) {
      rc = sip_fail(SIPNET_GPU_ERR_UNKNOWN_EVENT, "reading event file: unknown event type %s", typeStr);
      break;
    }
    if ((type == SIPNET_EV_LEAFON || type == SIPNET_EV_LEAFOFF) && !leafChecked) {
      /* checkForCalculatedLeafEvents(), events.c:251-261 */
      if (f->gdd || f->soilPhenol || params[SIPNET_P_leafOnDay] > 0 || params[SIPNET_P_leafOffDay] > 0) {
        rc = sip_fail(SIPNET_GPU_ERR_BAD_PARAMETER_VALUE,
                      "calculated leaf events (via leafOnDay/leafOffDay params or gdd/soil-phenol command-line options) "
                      "are not compatible with user-specified leaf events in event file");
        break;
      }
      leafChecked = 1;
    }
    if (s->nevents > 0 && ((year < lastYear) || ((year == lastYear) && (day < lastDay)))) { /* events.c:351-357 */
      rc = sip_fail(SIPNET_GPU_ERR_INPUT_FILE,
                    "reading event file: last event was at (%d, %d), next event is at (%d, %d); event records must be in "
                    "time-ascending order",
                    lastYear, lastDay, year, day);
      break;
    }
    if (s->nevents == cap) {
      cap = cap ? cap * 2 : 256;
      sipnet_gpu_event *n = (sipnet_gpu_event *)realloc(s->events, (size_t)cap * sizeof *n);
      if (!n) {
        rc = sip_fail(SIPNET_GPU_ERR_INTERNAL, "memory allocation failure in file processing");
        break;
      }
      s->events = n;
    }
    rc = parse_event_params(year, day, type, line + used, &s->events[s->nevents]);
    if (rc) break;
    s->nevents++;
    lastYear = year;
    lastDay = day;
  }
  fclose(in);
  return rc;
}

void sip_site_free(sip_site_data *s) {
  free(s->year);
  free(s->day);
  free(s->time);
  free(s->length);
  free(s->tair);
  free(s->tsoil);
  free(s->par);
  free(s->precip);
  free(s->vpd);
  free(s->vpdSoil);
  free(s->vPress);
  free(s->wspd);
  free(s->gdd);
  free(s->events);
  memset(s, 0, sizeof *s);
}

void sip_site_view(const sip_site_data *s, sipnet_gpu_site *v) {
  memset(v, 0, sizeof *v);
  v->nsteps = s->nsteps;
  v->year = s->year;
  v->day = s->day;
  v->time = s->time;
  v->length = s->length;
  v->tair = s->tair;
  v->tsoil = s->tsoil;
  v->par = s->par;
  v->precip = s->precip;
  v->vpd = s->vpd;
  v->vpdSoil = s->vpdSoil;
  v->vPress = s->vPress;
  v->wspd = s->wspd;
  v->gdd = s->gdd;
  v->nevents = s->nevents;
  v->events = s->events;
}
