/* tmb_io.c - the data formats either side of the path, host side, plain C99:
 * ILDG gauge configurations and SciDAC/ETMC propagator files in LIME containers, with the reference's
 * entry points (io/gauge.h:32-35, io/spinor.h:28) so that tmLQCD_read_gauge and a propagator dump work
 * without the reference's io/ directory.
 *
 * The reference delegates the container to the third-party c-lime library (usqcd-software/c-lime, "tested
 * with 1.2.3", README:10), which is neither vendored in the reference nor present in this image.  The
 * container layer below is therefore written from the published LIME record format: a 144-byte header
 *   u32 BE magic 0x456789AB | u16 BE version 1 | u16 BE flags (bit 15 MB, bit 14 ME) | u64 BE data length |
 *   128-byte NUL-padded type string
 * followed by the data, zero-padded to a multiple of 8 bytes.  Everything inside the records restates the
 * reference: record sequence and XML of io/gauge_write.c:22-58, io/utils_write_{xlf,ildg_format,checksum}.c,
 * payload order and endianness of io/gauge_{read,write}_binary.c (sites t,z,y,x slowest to fastest, links
 * x,y,z,t, big-endian IEEE), io/spinor_{read,write}_binary.c, the SciDAC checksum of io/dml.c:49-60 with the
 * zlib CRC-32 of io/DML_crc32.c, and the checks and return codes of io/gauge_read.c:29-194, io/spinor_read.c.
 * No arithmetic on fields happens here; nothing in this file touches the GPU.
 */
#include <complex.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>
#include "../../include/tmlqcd_b200.h"
#include "../../include/tmlqcd_b200_dropin.h"

double g_beta = 0., g_rgi_C1 = 0.; /* global.h:199,213: only printed into xlf-info */
int g_disable_IO_checks = 0;       /* global.h:77 */
int gauge_precision_read_flag = 64; /* read_input.h:69 */
tmb_gauge_info GaugeInfo = {0., 0, {0, 0}, NULL, NULL};

/* ------------------------------------------------------------------ big-endian conversion (io/utils.c) */
static inline uint64_t bswap64(uint64_t x) { return __builtin_bswap64(x); }
static inline uint32_t bswap32(uint32_t x) { return __builtin_bswap32(x); }
static int host_is_little(void) { const uint16_t one = 1; return *(const unsigned char *)&one == 1; }
/* All accesses go through memcpy: the buffers are su3 / double / unsigned char objects, and reading them through
 * uint64_t / uint32_t lvalues would break the aliasing rules (gcc -O2 then drops the stores that filled them). */
static void be64_copy(void *dst, const void *src, size_t n) { /* n doubles */
  if (!host_is_little()) { memcpy(dst, src, 8 * n); return; }
  const unsigned char *s = (const unsigned char *)src; unsigned char *d = (unsigned char *)dst;
  for (size_t i = 0; i < n; i++) { uint64_t u; memcpy(&u, s + 8 * i, 8); u = bswap64(u); memcpy(d + 8 * i, &u, 8); }
}
static void be32_from_double(void *dst, const double *src, size_t n) {
  unsigned char *d = (unsigned char *)dst;
  for (size_t i = 0; i < n; i++) { float f = (float)src[i]; uint32_t u; memcpy(&u, &f, 4); if (host_is_little()) u = bswap32(u); memcpy(d + 4 * i, &u, 4); }
}
static void double_from_be32(double *dst, const void *src, size_t n) {
  const unsigned char *s = (const unsigned char *)src;
  for (size_t i = 0; i < n; i++) { uint32_t u; memcpy(&u, s + 4 * i, 4); if (host_is_little()) u = bswap32(u); float f; memcpy(&f, &u, 4); dst[i] = (double)f; }
}

/* ------------------------------------------------------------------ SciDAC checksum: io/dml.c:49-60, io/DML_crc32.c */
static uint32_t crc_table[256];
static void crc_init(void) {
  static int done = 0;
  if (done) return;
  for (uint32_t n = 0; n < 256; n++) { uint32_t c = n; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1; crc_table[n] = c; }
  done = 1;
}
static uint32_t crc32_buf(const unsigned char *buf, size_t len) {
  uint32_t c = 0xffffffffu;
  for (size_t i = 0; i < len; i++) c = crc_table[(c ^ buf[i]) & 0xff] ^ (c >> 8);
  return c ^ 0xffffffffu;
}
static inline uint32_t rotl32(uint32_t x, unsigned r) { return r ? (x << r) | (x >> (32 - r)) : x; }
static void checksum_accum(DML_Checksum *cs, uint64_t rank, const void *buf, size_t size) {
  const uint32_t work = crc32_buf((const unsigned char *)buf, size);
  cs->suma ^= rotl32(work, (unsigned)(rank % 29));
  cs->sumb ^= rotl32(work, (unsigned)(rank % 31));
}

/* ------------------------------------------------------------------ LIME container */
#define LIME_HDR 144
#define LIME_MAGIC 0x456789abu
static void put_be(unsigned char *p, uint64_t v, int n) { for (int i = 0; i < n; i++) p[i] = (unsigned char)(v >> (8 * (n - 1 - i))); }
static uint64_t get_be(const unsigned char *p, int n) { uint64_t v = 0; for (int i = 0; i < n; i++) v = (v << 8) | p[i]; return v; }
static int lime_write_header(FILE *fp, int MB, int ME, const char *type, uint64_t bytes) {
  unsigned char b[LIME_HDR];
  memset(b, 0, sizeof(b));
  put_be(b, LIME_MAGIC, 4); put_be(b + 4, 1, 2); put_be(b + 6, (uint64_t)((MB ? 0x8000 : 0) | (ME ? 0x4000 : 0)), 2);
  put_be(b + 8, bytes, 8);
  strncpy((char *)b + 16, type, 127);
  return fwrite(b, 1, LIME_HDR, fp) == LIME_HDR ? 0 : -1;
}
static int lime_pad(FILE *fp, uint64_t bytes) {
  static const unsigned char z[8] = {0};
  const size_t p = (size_t)((8 - bytes % 8) % 8);
  return fwrite(z, 1, p, fp) == p ? 0 : -1;
}
static int lime_write_record(FILE *fp, int MB, int ME, const char *type, const void *data, uint64_t bytes) {
  if (lime_write_header(fp, MB, ME, type, bytes)) return -1;
  if (fwrite(data, 1, bytes, fp) != bytes) return -1;
  return lime_pad(fp, bytes);
}
typedef struct { char type[129]; uint64_t bytes; long data_pos; } lime_rec;
/* next record header; 0 ok, 1 end of file, -1 malformed.  Leaves the file position at the record's data. */
static int lime_next(FILE *fp, lime_rec *r, const lime_rec *prev) {
  unsigned char b[LIME_HDR];
  if (prev && fseek(fp, prev->data_pos + (long)(prev->bytes + (8 - prev->bytes % 8) % 8), SEEK_SET) != 0) return -1;
  const size_t got = fread(b, 1, LIME_HDR, fp);
  if (got == 0) return 1;
  if (got != LIME_HDR || get_be(b, 4) != LIME_MAGIC) return -1;
  r->bytes = get_be(b + 8, 8);
  memcpy(r->type, b + 16, 128); r->type[128] = 0;
  r->data_pos = ftell(fp);
  /* a length field that points past the end of the file is a damaged header: refuse it here, before anybody
   * allocates or seeks by it */
  if (r->data_pos < 0 || fseek(fp, 0, SEEK_END) != 0) return -1;
  const long size = ftell(fp);
  if (fseek(fp, r->data_pos, SEEK_SET) != 0 || size < r->data_pos || r->bytes > (uint64_t)(size - r->data_pos)) return -1;
  return 0;
}
static char *lime_read_message(FILE *fp, const lime_rec *r) {
  char *buf = (char *)calloc(r->bytes + 1, 1);
  if (!buf) return NULL;
  if (fread(buf, 1, r->bytes, fp) != r->bytes) { free(buf); return NULL; }
  return buf;
}

/* ------------------------------------------------------------------ XML records */
static int parse_checksum_xml(char *message, DML_Checksum *cs) { /* io/utils_parse_checksum_xml.c */
  int a = 0, b = 0;
  for (char *pos = strtok(message, "<> \n\t"); pos; pos = strtok(NULL, "<> \n\t")) {
    if (!strncmp(pos, "suma", 4)) { pos = strtok(NULL, "<> \n\t"); if (!pos) break; sscanf(pos, "%x", &cs->suma); a = 1; }
    else if (!strncmp(pos, "sumb", 4)) { pos = strtok(NULL, "<> \n\t"); if (!pos) break; sscanf(pos, "%x", &cs->sumb); b = 1; }
  }
  return a && b;
}
typedef struct { int lx, ly, lz, lt, prec; } ildg_format;
static int parse_ildgformat_xml(char *message, ildg_format *f) { /* io/utils_parse_ildgformat_xml.c */
  int n = 0;
  for (char *pos = strtok(message, "<> \n\t"); pos; pos = strtok(NULL, "<> \n\t")) {
    int *dst = NULL;
    if (!strncmp(pos, "precision", 9)) dst = &f->prec;
    else if (!strncmp(pos, "lx", 2)) dst = &f->lx;
    else if (!strncmp(pos, "ly", 2)) dst = &f->ly;
    else if (!strncmp(pos, "lz", 2)) dst = &f->lz;
    else if (!strncmp(pos, "lt", 2)) dst = &f->lt;
    if (dst) { pos = strtok(NULL, "<> \n\t"); if (!pos) break; if (sscanf(pos, "%d", dst) == 1) n++; }
  }
  return n >= 5;
}
static int write_checksum_record(FILE *fp, const DML_Checksum *cs, const char *name) { /* io/utils_write_checksum.c:22-49 */
  char m[512];
  snprintf(m, sizeof(m), "<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n<scidacChecksum>\n  <version>1.0</version>\n"
                         "  <suma>%08x</suma>\n  <sumb>%08x</sumb>\n</scidacChecksum>", cs->suma, cs->sumb);
  return lime_write_record(fp, 0, 1, name ? name : "scidac-checksum", m, strlen(m));
}

/* ------------------------------------------------------------------ gauge configurations */
/* io/params_construct_xlfInfo.c */
paramsXlfInfo *construct_paramsXlfInfo(double const plaq, int const counter) {
  struct timeval t1;
  paramsXlfInfo *info = (paramsXlfInfo *)calloc(1, sizeof(paramsXlfInfo));
  if (!info) { fprintf(stderr, "Could not allocate paramsXlfInfo.\n"); exit(500); }
  gettimeofday(&t1, NULL);
  info->plaq = plaq; info->counter = counter;
  info->beta = g_beta; info->kappa = g_kappa; info->mu = g_mu / 2. / g_kappa; info->c2_rec = g_rgi_C1;
  info->time = t1.tv_sec;
  strcpy(info->package_version, "tmlqcd-b200");
  info->mubar = g_mubar / 2. / g_kappa; info->epsilonbar = g_epsbar / 2. / g_kappa;
  strncpy(info->date, ctime(&t1.tv_sec), sizeof(info->date) - 1);
  return info;
}

/* io/gauge_write.c:22-58 with io/gauge_write_binary.c:125-215 (single process) */
int write_gauge_field(char *filename, int prec, paramsXlfInfo const *xlf) {
  if (prec != 64 && prec != 32) { fprintf(stderr, "write_gauge_field: precision must be 64 or 32\n"); return -1; }
  if (!g_gauge_field) { fprintf(stderr, "write_gauge_field: no gauge field (tmb_dropin_init first)\n"); return -1; }
  if (VOLUME != T * LX * LY * LZ) { fprintf(stderr, "%s: inconsistent lattice globals (VOLUME = %d, T LX LY LZ = %d %d %d %d).\n", __func__, VOLUME, T, LX, LY, LZ); return -1; }
  FILE *fp = fopen(filename, "w");
  if (!fp) { fprintf(stderr, "Failed to create writer. Aborting...\n"); return -1; }
  crc_init();
  char m[1024];
  int st = 0;
  if (xlf) { /* io/utils_write_xlf.c:22-64 */
    if (xlf->kappa != 0.0)
      snprintf(m, sizeof(m), "plaquette = %14.12f\n trajectory nr = %d\n beta = %.12f, kappa = %.12f, mu = %.12f, c2_rec = %f\n"
                             " time = %ld\n hmcversion = %s\n mubar = %.12f\n epsilonbar = %.12f\n date = %s",
               xlf->plaq, xlf->counter, xlf->beta, xlf->kappa, xlf->mu, xlf->c2_rec, xlf->time, xlf->package_version, xlf->mubar,
               xlf->epsilonbar, xlf->date);
    else
      snprintf(m, sizeof(m), "plaquette = %e\n trajectory nr = %d\n beta = %.12f\n kappa = %.12f\n 2*kappa*mu = %.12f\n c2_rec = %f\n date = %s",
               xlf->plaq, xlf->counter, xlf->beta, xlf->kappa, xlf->mu, xlf->c2_rec, xlf->date);
    st |= lime_write_record(fp, 1, 1, "xlf-info", m, strlen(m));
  }
  snprintf(m, sizeof(m), "<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n<ildgFormat xmlns=\"http://www.lqcd.org/ildg\"\n"
                         "            xmlns:xsi=\"http://www.w3.org/2001/XMLSchema-instance\"\n"
                         "            xsi:schemaLocation=\"http://www.lqcd.org/ildg/filefmt.xsd\">\n"
                         "  <version>1.0</version>\n  <field>su3gauge</field>\n  <precision>%d</precision>\n"
                         "  <lx>%d</lx>\n  <ly>%d</ly>\n  <lz>%d</lz>\n  <lt>%d</lt>\n</ildgFormat>", prec, LX, LY, LZ, T);
  st |= lime_write_record(fp, 1, 0, "ildg-format", m, strlen(m)); /* io/utils_write_ildg_format.c */
  const size_t site_bytes = (size_t)4 * sizeof(su3) * (size_t)prec / 64;
  const uint64_t bytes = (uint64_t)VOLUME * site_bytes;
  st |= lime_write_header(fp, 0, 0, "ildg-binary-data", bytes);
  DML_Checksum cs = {0, 0};
  unsigned char buf[4 * sizeof(su3)];
  su3 tmp[4];
  memset(tmp, 0, sizeof(tmp));
  static const int file_mu[4] = {1, 2, 3, 0}; /* file order x,y,z,t; memory order t,x,y,z */
  for (int t = 0; t < T; t++) for (int z = 0; z < LZ; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
    const uint64_t rank = (uint64_t)(((t * LZ + z) * LY + y) * LX + x);
    const int ix = ((t * LX + x) * LY + y) * LZ + z; /* g_ipt[t][x][y][z], geometry_eo.c:290 */
    for (int k = 0; k < 4; k++) tmp[k] = g_gauge_field[ix][file_mu[k]];
    if (prec == 64) be64_copy(buf, tmp, 72); else be32_from_double(buf, (const double *)tmp, 72);
    checksum_accum(&cs, rank, buf, site_bytes);
    if (fwrite(buf, 1, site_bytes, fp) != site_bytes) st = -1;
  }
  st |= lime_pad(fp, bytes);
  st |= write_checksum_record(fp, &cs, NULL);
  if (g_debug_level > 0 && g_proc_id == 0) {
    printf("# Scidac checksums for gaugefield %s:\n#   Calculated            : A = %#010x B = %#010x.\n", filename, cs.suma, cs.sumb);
    fflush(stdout);
  }
  if (fclose(fp)) st = -1;
  return st ? -2 : 0;
}

/* io/gauge_read.c:29-194 with io/gauge_read_binary.c:125-215 */
int read_gauge_field(char *filename, su3 **const gf) {
  FILE *fp = fopen(filename, "r");
  if (!fp) {
    fprintf(stderr, "\nUnable to open file for reading.\nPlease verify file existence and access rights.\nUnable to continue.\n");
    return -1;
  }
  crc_init();
  if (VOLUME != T * LX * LY * LZ) { /* the site loop below indexes gf with T, LX, LY, LZ and sizes the record with VOLUME */
    fprintf(stderr, "read_gauge_field: inconsistent lattice globals (VOLUME = %d, T LX LY LZ = %d %d %d %d).\n", VOLUME, T, LX, LY, LZ);
    fclose(fp); return -1;
  }
  lime_rec r, prev;
  int have_prev = 0, status, gauge_read = 0, dml_read = 0, fmt_read = 0;
  DML_Checksum calc = {0, 0}, stored = {0, 0};
  ildg_format fmt = {0, 0, 0, 0, 0};
  const int want_prec = gauge_precision_read_flag;
  GaugeInfo.gaugeRead = 0;
  if (g_proc_id == 0 && g_disable_IO_checks) fprintf(stdout, "# WARNING: IO CHECKS HAVE BEEN DISABLED\n");
  while ((status = lime_next(fp, &r, have_prev ? &prev : NULL)) == 0) {
    prev = r; have_prev = 1;
    if (!strcmp("ildg-binary-data", r.type)) {
      if (gauge_read && !g_disable_IO_checks) {
        fprintf(stderr, "In gauge file %s, multiple LIME records with name: \"ildg-binary-data\" found.\n", filename);
        fprintf(stderr, "Unable to verify integrity of the gauge field data.\n");
        fclose(fp); return -1;
      }
      const uint64_t full = (uint64_t)VOLUME * 4 * sizeof(su3);
      if (r.bytes != full / (want_prec == 64 ? 1 : 2)) {
        fprintf(stderr, "Lattice size and precision found in data file do not match those requested at input.\n");
        fprintf(stderr, "Expected LX = %d, LY = %d, LZ = %d, LT = %d, and %s precision.\n", LX, LY, LZ, T, want_prec == 64 ? "double" : "single");
        fprintf(stderr, "Expected %lu bytes, found %lu bytes.\n", (unsigned long)(full / (want_prec == 64 ? 1 : 2)), (unsigned long)r.bytes);
        fprintf(stderr, "Check input parameters T, L (LX, LY, LZ) and GaugeConfigReadPrecision.\n");
        fprintf(stderr, "Gauge file reading failed at binary part, unable to proceed.\n");
        fclose(fp); return -1;
      }
      const size_t site_bytes = (size_t)4 * sizeof(su3) * (size_t)want_prec / 64;
      unsigned char buf[4 * sizeof(su3)];
      su3 tmp[4];
      calc.suma = calc.sumb = 0;
      for (int t = 0; t < T; t++) for (int z = 0; z < LZ; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
        if (fread(buf, 1, site_bytes, fp) != site_bytes) {
          fprintf(stderr, "LIME read error occurred while reading in gauge_read_binary!\n");
          fprintf(stderr, "Gauge file reading failed at binary part, unable to proceed.\n");
          fclose(fp); return -1;
        }
        checksum_accum(&calc, (uint64_t)(((t * LZ + z) * LY + y) * LX + x), buf, site_bytes);
        if (want_prec == 64) be64_copy(tmp, buf, 72); else double_from_be32((double *)tmp, buf, 72);
        const int ix = ((t * LX + x) * LY + y) * LZ + z;
        gf[ix][1] = tmp[0]; gf[ix][2] = tmp[1]; gf[ix][3] = tmp[2]; gf[ix][0] = tmp[3];
      }
      gauge_read = 1; GaugeInfo.gaugeRead = 1; GaugeInfo.checksum = calc;
    } else if (!strcmp("scidac-checksum", r.type)) {
      if (dml_read && !g_disable_IO_checks) {
        fprintf(stderr, "In gauge file %s, multiple LIME records with name: \"scidac-checksum\" found.\n", filename);
        fprintf(stderr, "Unable to verify integrity of the gauge field data.\n");
        fclose(fp); return -1;
      }
      char *m = lime_read_message(fp, &r);
      if (m) { dml_read = parse_checksum_xml(m, &stored); free(m); }
    } else if (!strcmp("xlf-info", r.type)) {
      free(GaugeInfo.xlfInfo); GaugeInfo.xlfInfo = lime_read_message(fp, &r);
    } else if (!strcmp("ildg-data-lfn", r.type)) {
      free(GaugeInfo.ildg_data_lfn); GaugeInfo.ildg_data_lfn = lime_read_message(fp, &r);
    } else if (!strcmp("ildg-format", r.type)) {
      if (fmt_read && !g_disable_IO_checks) {
        fprintf(stderr, "In gauge file %s, multiple LIME records with name: \"ildg-format\" found.\n", filename);
        fprintf(stderr, "Unable to verify integrity of the gauge field data.\n");
        fclose(fp); return -1;
      }
      char *m = lime_read_message(fp, &r);
      if (m) { fmt_read = parse_ildgformat_xml(m, &fmt); free(m); }
    }
  }
  fclose(fp);
  if (status < 0) fprintf(stderr, "ReaderNextRecord returned status %d.\n", status);
  if (!g_disable_IO_checks) {
    if (!fmt_read) {
      fprintf(stderr, "LIME record with name: \"ildg-format\", in gauge file %s either missing or malformed.\n", filename);
      fprintf(stderr, "Unable to verify gauge field size or precision.\n");
      return -1;
    }
    if (!gauge_read) {
      fprintf(stderr, "LIME record with name: \"ildg-binary-data\", in gauge file %s either missing or malformed.\n", filename);
      fprintf(stderr, "No gauge field was read, unable to proceed.\n");
      return -1;
    }
    if (!dml_read) {
      fprintf(stderr, "LIME record with name: \"scidac-checksum\", in gauge file %s either missing or malformed.\n", filename);
      fprintf(stderr, "Unable to verify integrity of gauge field data.\n");
      return -1;
    }
    if (g_proc_id == 0 && g_debug_level > 0) {
      printf("# Scidac checksums for gaugefield %s:\n", filename);
      printf("#   Calculated            : A = %#010x B = %#010x.\n", calc.suma, calc.sumb);
      printf("#   Read from LIME headers: A = %#010x B = %#010x.\n", stored.suma, stored.sumb);
      fflush(stdout);
    }
    if (calc.suma != stored.suma) {
      fprintf(stderr, "For gauge file %s, calculated and stored values for SciDAC checksum A do not match.\n", filename);
      return -1;
    }
    if (calc.sumb != stored.sumb) {
      fprintf(stderr, "For gauge file %s, calculated and stored values for SciDAC checksum B do not match.\n", filename);
      return -1;
    }
    if (g_proc_id == 0 && g_debug_level > 0) { /* gauge_read.c:172-180 */
      printf("# Reading ildg-format record:\n#   Precision = %d bits (%s).\n", fmt.prec, fmt.prec == 64 ? "double" : "single");
      printf("#   Lattice size: LX = %d, LY = %d, LZ = %d, LT = %d.\n", fmt.lx, fmt.ly, fmt.lz, fmt.lt);
      printf("# Input parameters:\n#   Precision = %d bits (%s).\n", want_prec, want_prec == 64 ? "double" : "single");
      printf("#   Lattice size: LX = %d, LY = %d, LZ = %d, LT = %d.\n", LX, LY, LZ, T * g_nproc_t);
    }
    /* stricter than the reference, which only prints the two (gauge_read.c:172-180): a file whose extents are a
     * permutation of the requested ones has the right size and the right checksum, and every link in the wrong place */
    if (fmt.lx != LX || fmt.ly != LY || fmt.lz != LZ || fmt.lt != T * g_nproc_t) {
      fprintf(stderr, "For gauge file %s, the ildg-format record (LX = %d, LY = %d, LZ = %d, LT = %d) does not match the input "
                      "parameters (LX = %d, LY = %d, LZ = %d, LT = %d).\n", filename, fmt.lx, fmt.ly, fmt.lz, fmt.lt, LX, LY, LZ, T * g_nproc_t);
      return -1;
    }
  } else if (!gauge_read) return -1;
  g_update_gauge_copy = 1; /* io/gauge_read.c:186: the device copy is stale */
  return 0;
}

/* ------------------------------------------------------------------ propagators
 * One flavour, PropInfo.format == 0: the records op_write_prop (operator.c:532-605) writes: propagator-type,
 * [xlf-info copy], gauge-scidac-checksum-copy, inverter-info (io/spinor_write_info.c, utils_write_inverter_info.c),
 * etmc-propagator-format (io/spinor_write_propagator_format.c, including its lz = lx quirk), scidac-binary-data
 * (io/spinor_write_binary.c: sites t,z,y,x, even sites from s, odd from r), scidac-checksum. */
int tmb_write_propagator(const char *filename, spinor *const s, spinor *const r, int prec, double epssq, int iter,
                         const char *solver_name, int append) {
  if (prec != 64 && prec != 32) { fprintf(stderr, "tmb_write_propagator: precision must be 64 or 32\n"); return -1; }
  if (VOLUME != T * LX * LY * LZ) { fprintf(stderr, "%s: inconsistent lattice globals (VOLUME = %d, T LX LY LZ = %d %d %d %d).\n", __func__, VOLUME, T, LX, LY, LZ); return -1; }
  FILE *fp = fopen(filename, append ? "a" : "w");
  if (!fp) { fprintf(stderr, "Failed to create writer. Aborting...\n"); return -1; }
  crc_init();
  int st = 0;
  char m[1024];
  struct timeval t1; gettimeofday(&t1, NULL);
  if (!append) {
    st |= lime_write_record(fp, 1, 1, "propagator-type", "DiracFermion_Sink", strlen("DiracFermion_Sink"));
    if (GaugeInfo.xlfInfo) st |= lime_write_record(fp, 1, 0, "xlf-info", GaugeInfo.xlfInfo, strlen(GaugeInfo.xlfInfo));
    st |= write_checksum_record(fp, &GaugeInfo.checksum, "gauge-scidac-checksum-copy");
    if (GaugeInfo.ildg_data_lfn)
      st |= lime_write_record(fp, 1, 1, "gauge-ildg-data-lfn-copy", GaugeInfo.ildg_data_lfn, strlen(GaugeInfo.ildg_data_lfn));
  }
  snprintf(m, sizeof(m), "solver = %s\nepssq = %e\nnoiter = %d\nkappa = %.12f, mu = %.12f\ninverter version = %s\ndate = %s",
           solver_name ? solver_name : "CG", epssq, iter, g_kappa, g_mu / 2. / g_kappa, "tmlqcd-b200", ctime(&t1.tv_sec));
  st |= lime_write_record(fp, 1, 0, "inverter-info", m, strlen(m));
  snprintf(m, sizeof(m), "<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n<etmcFormat>\n  <field>diracFermion</field>\n"
                         "  <precision>%d</precision>\n  <flavours>%d</flavours>\n  <lx>%d</lx>\n  <ly>%d</ly>\n  <lz>%d</lz>\n"
                         "  <lt>%d</lt>\n</etmcFormat>", prec, 1, LX, LY, LX /* sic: spinor_write_propagator_format.c:38 */, T);
  st |= lime_write_record(fp, 0, 1, "etmc-propagator-format", m, strlen(m));
  const size_t site_bytes = sizeof(spinor) * (size_t)prec / 64;
  const uint64_t bytes = (uint64_t)VOLUME * site_bytes;
  st |= lime_write_header(fp, 1, 0, "scidac-binary-data", bytes);
  DML_Checksum cs = {0, 0};
  unsigned char buf[sizeof(spinor)];
  int ne = 0, no = 0;
  /* g_lexic2eosub is the rank among the sites of the same parity in lexicographic order (geometry_eo.c:869-884);
   * the file order t,z,y,x is not lexicographic (t,x,y,z), so build the table once */
  int *eosub = (int *)malloc(sizeof(int) * (size_t)VOLUME);
  if (!eosub) { fclose(fp); return -1; }
  for (int ix = 0; ix < VOLUME; ix++) {
    const int z = ix % LZ, y = (ix / LZ) % LY, x = (ix / (LZ * LY)) % LX, t = ix / (LZ * LY * LX);
    eosub[ix] = ((t + x + y + z) % 2 == 0) ? ne++ : no++;
  }
  for (int t = 0; t < T; t++) for (int z = 0; z < LZ; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
    const uint64_t rank = (uint64_t)(((t * LZ + z) * LY + y) * LX + x);
    const int ix = ((t * LX + x) * LY + y) * LZ + z;
    const spinor *p = ((t + x + y + z) % 2 == 0 ? s : r) + eosub[ix];
    if (prec == 64) be64_copy(buf, p, 24); else be32_from_double(buf, (const double *)p, 24);
    checksum_accum(&cs, rank, buf, site_bytes);
    if (fwrite(buf, 1, site_bytes, fp) != site_bytes) st = -1;
  }
  free(eosub);
  st |= lime_pad(fp, bytes);
  st |= write_checksum_record(fp, &cs, NULL);
  if (fclose(fp)) st = -1;
  return st ? -2 : 0;
}

/* io/spinor_read.c:27-150 (r != NULL: even/odd pair), position-th scidac-binary-data record */
int read_spinor(spinor *const s, spinor *const r, char *filename, const int position_) {
  if (VOLUME != T * LX * LY * LZ) { fprintf(stderr, "%s: inconsistent lattice globals (VOLUME = %d, T LX LY LZ = %d %d %d %d).\n", __func__, VOLUME, T, LX, LY, LZ); return -1; }
  FILE *fp = fopen(filename, "r");
  if (!fp) {
    fprintf(stderr, "\nUnable to open file for reading.\nPlease verify file existence and access rights.\nUnable to continue.\n");
    return -1;
  }
  crc_init();
  lime_rec rec, prev;
  int have_prev = 0, status, position = position_, getpos = 0, found = 0;
  /* propagator type: DiracFermion_Source_Sink_Pairs stores source and sink alternately (spinor_read.c:40-44) */
  while ((status = lime_next(fp, &rec, have_prev ? &prev : NULL)) == 0) {
    prev = rec; have_prev = 1;
    if (!strcmp("propagator-type", rec.type)) {
      char *m = lime_read_message(fp, &rec);
      if (m && !strcmp(m, "DiracFermion_Source_Sink_Pairs")) position = 2 * position_ + 1;
      else if (m && (!strcmp(m, "DiracFermion_ScalarSource_TwelveSink") || !strcmp(m, "DiracFermion_ScalarSource_FourSink"))) { free(m); fclose(fp); return -2; }
      free(m);
      break;
    }
    if (!strcmp("source-type", rec.type)) break;
  }
  rewind(fp); have_prev = 0;
  while ((status = lime_next(fp, &rec, have_prev ? &prev : NULL)) == 0) {
    prev = rec; have_prev = 1;
    if (!strcmp("scidac-binary-data", rec.type)) { if (getpos == position) { found = 1; break; } ++getpos; }
  }
  if (!found) {
    fprintf(stderr, "Unable to find requested LIME record scidac-binary-data in file %s.\nEnd of file reached before record was found.\n", filename);
    fclose(fp); return -5;
  }
  int prec;
  if (rec.bytes == (uint64_t)VOLUME * sizeof(spinor)) prec = 64;
  else if (rec.bytes == (uint64_t)VOLUME * sizeof(spinor) / 2) prec = 32;
  else {
    fprintf(stderr, "Length of scidac-binary-data record in %s does not match input parameters.\n", filename);
    fprintf(stderr, "Found %lu bytes.\n", (unsigned long)rec.bytes);
    fclose(fp); return -6;
  }
  if (g_proc_id == 0 && g_debug_level >= 0) printf("# %s precision read (%d bits).\n", prec == 64 ? "Double" : "Single", prec);
  const size_t site_bytes = sizeof(spinor) * (size_t)prec / 64;
  unsigned char buf[sizeof(spinor)];
  DML_Checksum calc = {0, 0}, stored = {0, 0};
  int *eosub = (int *)malloc(sizeof(int) * (size_t)VOLUME), ne = 0, no = 0;
  if (!eosub) { fclose(fp); return -7; }
  for (int ix = 0; ix < VOLUME; ix++) {
    const int z = ix % LZ, y = (ix / LZ) % LY, x = (ix / (LZ * LY)) % LX, t = ix / (LZ * LY * LX);
    eosub[ix] = ((t + x + y + z) % 2 == 0) ? ne++ : no++;
  }
  for (int t = 0; t < T; t++) for (int z = 0; z < LZ; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
    if (fread(buf, 1, site_bytes, fp) != site_bytes) { free(eosub); fclose(fp); fprintf(stderr, "read_binary_spinor_data failed\n"); return -7; }
    checksum_accum(&calc, (uint64_t)(((t * LZ + z) * LY + y) * LX + x), buf, site_bytes);
    const int ix = ((t * LX + x) * LY + y) * LZ + z;
    spinor *p = ((t + x + y + z) % 2 == 0 ? s : r) + eosub[ix];
    if (prec == 64) be64_copy(p, buf, 24); else double_from_be32((double *)p, buf, 24);
  }
  free(eosub);
  int dml = 0;
  prev = rec;
  while ((status = lime_next(fp, &rec, &prev)) == 0) {
    prev = rec;
    if (!strcmp("scidac-checksum", rec.type)) { char *m = lime_read_message(fp, &rec); if (m) { dml = parse_checksum_xml(m, &stored); free(m); } break; }
    if (!strcmp("scidac-binary-data", rec.type) || !strcmp("ildg-binary-data", rec.type)) break;
  }
  fclose(fp);
  if (!dml) {
    fprintf(stderr, "LIME record with name: \"scidac-checksum\", in gauge file %s either missing or malformed.\n", filename);
    fprintf(stderr, "Unable to verify integrity of gauge field data.\n");
    return -1;
  }
  if (g_proc_id == 0 && g_debug_level >= 0) {
    printf("# Scidac checksums for DiracFermion field %s position %d:\n", filename, position);
    printf("#   Calculated            : A = %#010x B = %#010x.\n", calc.suma, calc.sumb);
    printf("#   Read from LIME headers: A = %#010x B = %#010x.\n", stored.suma, stored.sumb);
  }
  /* the reference prints the two and returns 0 either way (spinor_read.c:141-148); same here, plus a warning */
  if (calc.suma != stored.suma || calc.sumb != stored.sumb)
    fprintf(stderr, "WARNING: SciDAC checksum of DiracFermion field %s position %d does not match the stored one.\n", filename, position);
  return 0;
}
