/*
 * sip_config.c -- run configuration: defaults, sipnet.in, command line, validation, config dump.
 *
 * Semantics follow the reference's Context (src/common/context.[ch]), its
 * sipnet.in reader (src/sipnet/frontend.c:35-128) and CLI (src/sipnet/cli.c):
 * precedence default < input file < command line < calculated; keys are matched
 * after dropping non-alphanumerics and lower-casing (context.c:76-90), with the
 * legacy FILE_NAME alias; unknown keys are ignored with an info line; RUNTYPE
 * must be "standard" (frontend.c:24-33).
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <errno.h>
#include <getopt.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include "sip_host.h"

static _Thread_local char g_err[1024]; /* per thread: the readers and writers run on worker threads in many-site launches */
const char *sip_host_error(void) { return g_err; }
int sip_fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
void sip_info(int quiet, const char *fmt, ...) {
  if (quiet) return;
  va_list ap;
  va_start(ap, fmt);
  fputs("[INFO   ] ", stdout);
  vprintf(fmt, ap);
  va_end(ap);
}

/* ---- the settings table (context.c:35-66) ---------------------------------------------------- */
typedef struct {
  const char *key;   /* normalised name used for matching and for sorting the dump */
  const char *print; /* name printed in <prefix>.config */
  const char *cli;   /* --cli-flag for int flags, NULL otherwise */
  int isChar;
  size_t offset;
} setting_t;

#define FLAG(field) offsetof(sip_context, flags.field)
#define CTXF(field) offsetof(sip_context, field)
static const setting_t kSettings[] = {
    {"events", "EVENTS", "events", 0, FLAG(events)},
    {"gdd", "GDD", "gdd", 0, FLAG(gdd)},
    {"growthresp", "GROWTH_RESP", "growth-resp", 0, FLAG(growthResp)},
    {"leafwater", "LEAF_WATER", "leaf-water", 0, FLAG(leafWater)},
    {"litterpool", "LITTER_POOL", "litter-pool", 0, FLAG(litterPool)},
    {"snow", "SNOW", "snow", 0, FLAG(snow)},
    {"soilphenol", "SOIL_PHENOL", "soil-phenol", 0, FLAG(soilPhenol)},
    {"waterhresp", "WATER_HRESP", "water-hresp", 0, FLAG(waterHResp)},
    {"nitrogencycle", "NITROGEN_CYCLE", "nitrogen-cycle", 0, FLAG(nitrogenCycle)},
    {"anaerobic", "ANAEROBIC", "anaerobic", 0, FLAG(anaerobic)},
    {"flooding", "FLOODING", "flooding", 0, FLAG(flooding)},
    {"carbonsaturation", "CARBON_SATURATION", "carbon-saturation", 0, FLAG(carbonSaturation)},
    {"domainoutput", "DO_MAIN_OUTPUT", "do-main-output", 0, CTXF(doMainOutput)},
    {"dosingleoutputs", "DO_SINGLE_OUTPUT", "do-single-outputs", 0, CTXF(doSingleOutputs)},
    {"dumpconfig", "DUMP_CONFIG", "dump-config", 0, CTXF(dumpConfig)},
    {"printheader", "PRINT_HEADER", "print-header", 0, CTXF(printHeader)},
    {"quiet", "QUIET", "quiet", 0, CTXF(quiet)},
    {"paramfile", "PARAM_FILE", NULL, 1, CTXF(paramFile)},
    {"climfile", "CLIM_FILE", NULL, 1, CTXF(climFile)},
    {"outfile", "OUT_FILE", NULL, 1, CTXF(outFile)},
    {"outconfigfile", "OUT_CONFIG_FILE", NULL, 1, CTXF(outConfigFile)},
    {"eventsprefix", "EVENTS_PREFIX", NULL, 1, CTXF(eventsPrefix)},
    {"inputfile", "INPUT_FILE", NULL, 1, CTXF(inputFile)},
    {"restartin", "RESTART_IN", NULL, 1, CTXF(restartIn)},
    {"restartout", "RESTART_OUT", NULL, 1, CTXF(restartOut)},
    {"debuglogprefix", "DEBUG_LOG_PREFIX", NULL, 1, CTXF(debugLogPrefix)},
    {"fileprefix", "FILE_PREFIX", NULL, 1, CTXF(filePrefix)},
};
enum { kNumSettings = (int)(sizeof kSettings / sizeof kSettings[0]), kNumFlagSettings = 17 };

static void normalise(const char *name, char *key, size_t cap) { /* nameToKey(), context.c:76-90 */
  size_t k = 0;
  for (size_t i = 0; name[i] && k + 1 < cap; ++i)
    if (isalnum((unsigned char)name[i])) key[k++] = (char)tolower((unsigned char)name[i]);
  key[k] = '\0';
  if (strcmp(key, "filename") == 0) strcpy(key, "fileprefix");
}
static int find_setting(const char *name) {
  char key[SIP_NAME_MAX];
  normalise(name, key, sizeof key);
  for (int i = 0; i < kNumSettings; ++i)
    if (strcmp(kSettings[i].key, key) == 0) return i;
  return -1;
}
static int *int_field(sip_context *c, int i) { return (int *)((char *)c + kSettings[i].offset); }
static char *char_field(sip_context *c, int i) { return (char *)c + kSettings[i].offset; }
static void set_int(sip_context *c, int i, int v, int src) { /* updateIntContext(): higher or equal source wins */
  if (c->source[i] <= src) {
    *int_field(c, i) = v;
    c->source[i] = src;
  }
}
static void set_char(sip_context *c, int i, const char *v, int src) {
  if (c->source[i] <= src) {
    strncpy(char_field(c, i), v, SIP_NAME_MAX - 1);
    char_field(c, i)[SIP_NAME_MAX - 1] = '\0';
    c->source[i] = src;
  }
}

void sip_context_init(sip_context *c) {
  memset(c, 0, sizeof *c);
  c->flags.events = 1;
  c->flags.gdd = 1;
  c->flags.snow = 1;
  c->flags.waterHResp = 1;
  c->doMainOutput = 1;
  c->printHeader = 1;
  strcpy(c->eventsPrefix, "events");
  strcpy(c->inputFile, "sipnet.in");
  strcpy(c->filePrefix, "sipnet");
}

/* ---- command line (cli.c:144-233) --------------------------------------------------------------- */
static void usage(const char *prog) {
  printf("Usage: %s [OPTIONS]\n\n", prog);
  printf("Run SIPNET on a CUDA device (B200) for one site with the configured options; a drop-in for the\n"
         "reference `sipnet` binary's per-timestep loop, optionally over a parameter ensemble.\n\n");
  printf("  -i, --input-file <path>      config file ('sipnet.in')\n"
         "  -f, --file-prefix <name>     prefix of the .param / .clim / .out files ('sipnet'); alias --file-name\n"
         "  -e, --events-prefix <name>   prefix of events .in / .out ('events')\n"
         "      --ensemble-params <file> one .param path per line: run them as one ensemble; member k writes\n"
         "                               <prefix>.out.<k> (this implementation's extension)\n"
         "      --site-list <file>       several sites in one launch (extension).  One site per line:\n"
         "                               <file-prefix> [<events-prefix> [<member-list>]]; defaults: events next to the\n"
         "                               site's files, one member from <file-prefix>.param; all sites share this run's flags\n"
         "      --devices <n>            GPUs to use for --ensemble-params / --site-list launches (0 = all visible,\n"
         "                               the default): whole sites per GPU, or an even share of one site's members\n"
         "      --validation-math        run the general (reference-shaped) kernel instead of the optimistic one\n"
         "  model flags (prefix with no- to turn off): --events --gdd --growth-resp --leaf-water --litter-pool --snow\n"
         "      --soil-phenol --water-hresp --nitrogen-cycle --anaerobic --flooding --carbon-saturation\n"
         "  output flags: --do-main-output --do-single-outputs --dump-config --print-header --quiet\n"
         "  -h, --help   -v, --version\n");
}

int sip_parse_cli(sip_context *c, int argc, char **argv) {
  enum { OPT_RESTART_IN = 1001, OPT_RESTART_OUT, OPT_DEBUG_LOG, OPT_ENSEMBLE, OPT_VALIDATION, OPT_SITE_LIST, OPT_DEVICES };
  struct option opts[2 * kNumFlagSettings + 16];
  char names[kNumFlagSettings][40];
  int flagValue = 0, n = 0;
  for (int i = 0; i < kNumFlagSettings; ++i) {
    snprintf(names[i], sizeof names[i], "no-%s", kSettings[i].cli);
    opts[n++] = (struct option){kSettings[i].cli, no_argument, &flagValue, 1};
    opts[n++] = (struct option){names[i], no_argument, &flagValue, 0};
  }
  opts[n++] = (struct option){"input-file", required_argument, 0, 'i'};
  opts[n++] = (struct option){"file-prefix", required_argument, 0, 'f'};
  opts[n++] = (struct option){"file-name", required_argument, 0, 'f'};
  opts[n++] = (struct option){"events-prefix", required_argument, 0, 'e'};
  opts[n++] = (struct option){"restart-in", required_argument, 0, OPT_RESTART_IN};
  opts[n++] = (struct option){"restart-out", required_argument, 0, OPT_RESTART_OUT};
  opts[n++] = (struct option){"debug-log", required_argument, 0, OPT_DEBUG_LOG};
  opts[n++] = (struct option){"ensemble-params", required_argument, 0, OPT_ENSEMBLE};
  opts[n++] = (struct option){"site-list", required_argument, 0, OPT_SITE_LIST};
  opts[n++] = (struct option){"devices", required_argument, 0, OPT_DEVICES};
  opts[n++] = (struct option){"validation-math", no_argument, 0, OPT_VALIDATION};
  opts[n++] = (struct option){"help", no_argument, 0, 'h'};
  opts[n++] = (struct option){"version", no_argument, 0, 'v'};
  opts[n++] = (struct option){0, 0, 0, 0};
  optind = 1;
  int idx = 0, ch;
  while ((ch = getopt_long(argc, argv, "he:f:i:v", opts, &idx)) != -1) {
    switch (ch) {
      case 0:
        set_int(c, idx / 2, flagValue, SIP_SRC_COMMAND_LINE);
        break;
      case 'f':
        if (strlen(optarg) > SIP_NAME_MAX - 10)
          return sip_fail(SIPNET_GPU_ERR_BAD_CLI, "file prefix '%s' exceeds maximum length of %d characters", optarg,
                          SIP_NAME_MAX - 10);
        set_char(c, find_setting("fileprefix"), optarg, SIP_SRC_COMMAND_LINE);
        break;
      case 'e':
        if (strlen(optarg) >= SIP_NAME_MAX) return sip_fail(SIPNET_GPU_ERR_BAD_CLI, "events-prefix value too long");
        set_char(c, find_setting("eventsprefix"), optarg, SIP_SRC_COMMAND_LINE);
        break;
      case 'i':
        if (strlen(optarg) >= SIP_NAME_MAX) return sip_fail(SIPNET_GPU_ERR_BAD_CLI, "input filename too long");
        set_char(c, find_setting("inputfile"), optarg, SIP_SRC_COMMAND_LINE);
        break;
      case OPT_RESTART_IN:
        set_char(c, find_setting("restartin"), optarg, SIP_SRC_COMMAND_LINE);
        break;
      case OPT_RESTART_OUT:
        set_char(c, find_setting("restartout"), optarg, SIP_SRC_COMMAND_LINE);
        break;
      case OPT_DEBUG_LOG:
        set_char(c, find_setting("debuglogprefix"), optarg, SIP_SRC_COMMAND_LINE);
        break;
      case OPT_ENSEMBLE:
        strncpy(c->ensembleParamList, optarg, SIP_NAME_MAX - 1);
        break;
      case OPT_SITE_LIST:
        strncpy(c->siteList, optarg, SIP_NAME_MAX - 1);
        break;
      case OPT_DEVICES: {
        char *end = NULL;
        const long v = strtol(optarg, &end, 10);
        if (end == optarg || *end != '\0' || v < 0 || v > 64)
          return sip_fail(SIPNET_GPU_ERR_BAD_CLI, "--devices takes a count between 0 (all) and 64, got '%s'", optarg);
        c->devices = (int32_t)v;
      } break;
      case OPT_VALIDATION:
        c->validationMath = 1;
        break;
      case 'h':
        usage(argv[0]);
        c->helpOrVersion = 1;
        return 0;
      case 'v':
        printf("SIPNET-GPU (B200) drop-in for SIPNET version 2.1.0\n");
        c->helpOrVersion = 2;
        return 0;
      default:
        usage(argv[0]);
        return sip_fail(SIPNET_GPU_ERR_BAD_CLI, "bad command line argument");
    }
  }
  return 0;
}

/* ---- sipnet.in (frontend.c:35-128) ------------------------------------------------------------------ */
static int only_blank_after_strip(char *line) { /* stripComment(), util.c:33-47, comment char '!' */
  char *bang = strpbrk(line, "!");
  if (bang) *bang = '\0';
  return strlen(line) == strspn(line, " \t\n\r");
}

int sip_read_input_file(sip_context *c) {
  if (c->filePrefix[0] == '\0') /* validateFilename(), context.c:180-193 */
    return sip_fail(SIPNET_GPU_ERR_BAD_PARAMETER_VALUE, "filePrefix must be set for SIPNET to run");
  if (strlen(c->filePrefix) > SIP_NAME_MAX - 10)
    return sip_fail(SIPNET_GPU_ERR_BAD_PARAMETER_VALUE, "filePrefix is too long; max length is %d characters",
                    SIP_NAME_MAX - 10);
  sip_info(c->quiet, "Reading config from file %s\n", c->inputFile);
  FILE *in = fopen(c->inputFile, "r");
  if (!in) {
    fprintf(stderr, "Error reading '%s': %s\n", c->inputFile, strerror(errno));
    return sip_fail(SIPNET_GPU_ERR_FILE_OPEN, "cannot open %s", c->inputFile);
  }
  char line[1024];
  int rc = 0;
  while (rc == 0 && fgets(line, sizeof line, in) != NULL) {
    if (only_blank_after_strip(line)) continue;
    char *name = strtok(line, " \t=:");
    char *value = strtok(NULL, " \t=:\n\r");
    if (!name) continue;
    if (strcasecmp(name, "runtype") == 0) { /* obsolete; must be "standard" (frontend.c:24-33) */
      if (!value || strcasecmp(value, "standard") != 0)
        rc = sip_fail(SIPNET_GPU_ERR_BAD_PARAMETER_VALUE,
                      "SIPNET only supports the standard runtype mode; please fix %s and re-run", c->inputFile);
      continue;
    }
    const int i = find_setting(name);
    if (i < 0) {
      sip_info(c->quiet, "ignoring input file parameter %s\n", name);
      continue;
    }
    if (!value) {
      rc = sip_fail(SIPNET_GPU_ERR_BAD_PARAMETER_VALUE, "Error in input file: No value given for input item %s", name);
      continue;
    }
    if (!kSettings[i].isChar) {
      char *end = NULL;
      const long v = strtol(value, &end, 0);
      if (end && *end != '\0')
        rc = sip_fail(SIPNET_GPU_ERR_BAD_PARAMETER_VALUE, "ERROR in input file: Invalid value for %s: %s", name, value);
      else
        set_int(c, i, (int)v, SIP_SRC_INPUT_FILE);
    } else if (strcmp(value, "none") == 0) {
      set_char(c, i, "", SIP_SRC_INPUT_FILE);
    } else if (strlen(value) >= SIP_NAME_MAX) {
      rc = sip_fail(SIPNET_GPU_ERR_BAD_PARAMETER_VALUE, "ERROR in input file: value '%s' exceeds maximum length for %s (%d)",
                    value, name, SIP_NAME_MAX);
    } else {
      set_char(c, i, value, SIP_SRC_INPUT_FILE);
    }
  }
  fclose(in);
  return rc;
}

int sip_validate_context(const sip_context *c) { /* validateContext(), context.c:195-223 */
  int bad = 0;
  g_err[0] = '\0';
  if (c->flags.soilPhenol && c->flags.gdd) {
    strncat(g_err, "soil-phenol and gdd may not both be turned on; ", sizeof g_err - strlen(g_err) - 1);
    bad = 1;
  }
  if (c->flags.nitrogenCycle && !(c->flags.litterPool && c->flags.anaerobic)) {
    strncat(g_err, "nitrogen-cycle requires both litter-pool and anaerobic to be turned on; ",
            sizeof g_err - strlen(g_err) - 1);
    bad = 1;
  }
  if (c->flags.anaerobic && !c->flags.waterHResp) {
    strncat(g_err, "anaerobic requires water-hresp to be turned on; ", sizeof g_err - strlen(g_err) - 1);
    bad = 1;
  }
  if (c->flags.carbonSaturation && !c->flags.litterPool) {
    strncat(g_err, "carbon-saturation requires litter-pool to be turned on; ", sizeof g_err - strlen(g_err) - 1);
    bad = 1;
  }
  return bad ? SIPNET_GPU_ERR_BAD_PARAMETER_VALUE : 0;
}

int sip_derive_file_names(sip_context *c) { /* frontend.c:164-209 */
  char buf[SIP_NAME_MAX + 16];
  snprintf(buf, sizeof buf, "%s.param", c->filePrefix);
  set_char(c, find_setting("paramfile"), buf, SIP_SRC_CALCULATED);
  snprintf(buf, sizeof buf, "%s.clim", c->filePrefix);
  set_char(c, find_setting("climfile"), buf, SIP_SRC_CALCULATED);
  if (c->flags.events) {
    if (strlen(c->eventsPrefix) > SIP_NAME_MAX - sizeof(".out"))
      return sip_fail(SIPNET_GPU_ERR_BAD_PARAMETER_VALUE, "events-prefix value %s is too long", c->eventsPrefix);
    snprintf(c->eventsInFile, SIP_NAME_MAX, "%s.in", c->eventsPrefix);
    snprintf(c->eventsOutFile, SIP_NAME_MAX, "%s.out", c->eventsPrefix);
  } else {
    c->eventsInFile[0] = c->eventsOutFile[0] = '\0';
  }
  if (c->doMainOutput) {
    snprintf(buf, sizeof buf, "%s.out", c->filePrefix);
    set_char(c, find_setting("outfile"), buf, SIP_SRC_CALCULATED);
  }
  if (c->dumpConfig) {
    snprintf(buf, sizeof buf, "%s.config", c->filePrefix);
    set_char(c, find_setting("outconfigfile"), buf, SIP_SRC_CALCULATED);
  }
  return 0;
}

static int by_key(const void *a, const void *b) {
  return strcmp(kSettings[*(const int *)a].key, kSettings[*(const int *)b].key);
}

int sip_print_config(const sip_context *c, FILE *out, const char *timestamp) { /* printConfig(), context.c:225-267 */
  static const char *kSource[] = {"DEFAULT", "INPUT_FILE", "COMMAND_LINE", "CALCULATED", "TEST"};
  int order[kNumSettings];
  unsigned width = 0;
  for (int i = 0; i < kNumSettings; ++i) {
    order[i] = i;
    if (kSettings[i].isChar && strlen(kSettings[i].print) > width) width = (unsigned)strlen(kSettings[i].print);
  }
  qsort(order, kNumSettings, sizeof order[0], by_key);
  if (c->printHeader) {
    fprintf(out, "Final config for SIPNET run at %s\n", timestamp);
    fprintf(out, "%21s %13s %*s\n", "Name", "Source", (int)width, "Value");
  }
  for (int k = 0; k < kNumSettings; ++k) {
    const int i = order[k];
    if (kSettings[i].isChar)
      fprintf(out, "%21s %13s %*s\n", kSettings[i].print, kSource[c->source[i]], (int)width,
              (const char *)c + kSettings[i].offset);
    else
      fprintf(out, "%21s %13s %*d\n", kSettings[i].print, kSource[c->source[i]], (int)width,
              *(const int *)((const char *)c + kSettings[i].offset));
  }
  return 0;
}
