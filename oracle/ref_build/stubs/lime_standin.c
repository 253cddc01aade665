/* lime_standin.c - minimal implementation of the c-lime calls declared in stubs/lime.h (see there).
 * TEST INFRASTRUCTURE ONLY: lets the reference's unmodified io/ sources run in oracle/_ref. */
#include <stdlib.h>
#include <string.h>
#include "lime.h"

#define HDR 144
#define MAGIC 0x456789abu
static void put_be(unsigned char *p, uint64_t v, int n) { for (int i = 0; i < n; i++) p[i] = (unsigned char)(v >> (8 * (n - 1 - i))); }
static uint64_t get_be(const unsigned char *p, int n) { uint64_t v = 0; for (int i = 0; i < n; i++) v = (v << 8) | p[i]; return v; }
static n_uint64_t pad8(n_uint64_t n) { return (8 - n % 8) % 8; }

LimeRecordHeader *limeCreateHeader(int MB_flag, int ME_flag, char *type, n_uint64_t reclen) {
  LimeRecordHeader *h = calloc(1, sizeof(*h));
  h->lime_version = 1; h->MB_flag = MB_flag; h->ME_flag = ME_flag; h->data_length = reclen;
  h->type = calloc(129, 1); strncpy(h->type, type, 128);
  return h;
}
void limeDestroyHeader(LimeRecordHeader *h) { if (h) { free(h->type); free(h); } }

LimeWriter *limeCreateWriter(FILE *fp) {
  LimeWriter *w = calloc(1, sizeof(*w));
  w->fp = fp; w->first_record = 1; w->header_nextP = 1;
  return w;
}
int limeDestroyWriter(LimeWriter *w) { if (w) { limeWriterCloseRecord(w); free(w); } return LIME_SUCCESS; }
int limeWriteRecordHeader(LimeRecordHeader *props, LimeWriter *w) {
  unsigned char b[HDR];
  if (!w || !props) return LIME_ERR_PARAM;
  limeWriterCloseRecord(w);
  memset(b, 0, HDR);
  put_be(b, MAGIC, 4); put_be(b + 4, 1, 2);
  put_be(b + 6, (uint64_t)((props->MB_flag ? 0x8000 : 0) | (props->ME_flag ? 0x4000 : 0)), 2);
  put_be(b + 8, props->data_length, 8);
  strncpy((char *)b + 16, props->type, 127);
  if (fwrite(b, 1, HDR, w->fp) != HDR) return LIME_ERR_WRITE;
  w->bytes_total = props->data_length; w->bytes_left = props->data_length; w->bytes_pad = pad8(props->data_length);
  w->header_nextP = 0;
  return LIME_SUCCESS;
}
int limeWriteRecordData(void *source, n_uint64_t *nbytes, LimeWriter *w) {
  n_uint64_t n = *nbytes;
  if (w->header_nextP) return LIME_ERR_HEADER_NEXT;
  if (n > w->bytes_left) n = w->bytes_left;
  if (fwrite(source, 1, n, w->fp) != n) return LIME_ERR_WRITE;
  *nbytes = n; w->bytes_left -= n;
  if (w->bytes_left == 0) return limeWriterCloseRecord(w);
  return LIME_SUCCESS;
}
int limeWriterCloseRecord(LimeWriter *w) {
  static const unsigned char z[8] = {0};
  if (!w || w->header_nextP) return LIME_SUCCESS;
  if (w->bytes_left) { /* short record: fill */
    while (w->bytes_left) { n_uint64_t k = w->bytes_left < 8 ? w->bytes_left : 8; fwrite(z, 1, k, w->fp); w->bytes_left -= k; }
  }
  if (w->bytes_pad) fwrite(z, 1, w->bytes_pad, w->fp);
  w->bytes_pad = 0; w->header_nextP = 1;
  fflush(w->fp);
  return LIME_SUCCESS;
}

LimeReader *limeCreateReader(FILE *fp) {
  LimeReader *r = calloc(1, sizeof(*r));
  r->fp = fp; r->first_read = 0; r->header_nextP = 1;
  r->curr_header = limeCreateHeader(0, 0, "", 0);
  return r;
}
void limeDestroyReader(LimeReader *r) { if (r) { limeDestroyHeader(r->curr_header); free(r); } }
int limeReaderCloseRecord(LimeReader *r) {
  if (!r->header_nextP) {
    if (fseeko(r->fp, (off_t)(r->rec_start + r->bytes_total + r->bytes_pad), SEEK_SET) != 0) return LIME_ERR_SEEK;
    r->header_nextP = 1;
  }
  return LIME_SUCCESS;
}
int limeReaderNextRecord(LimeReader *r) {
  unsigned char b[HDR];
  int st = limeReaderCloseRecord(r);
  if (st != LIME_SUCCESS) return st;
  size_t got = fread(b, 1, HDR, r->fp);
  if (got == 0) return LIME_EOF;
  if (got != HDR || get_be(b, 4) != MAGIC) return LIME_ERR_READ;
  const unsigned flags = (unsigned)get_be(b + 6, 2);
  r->curr_header->MB_flag = (flags >> 15) & 1; r->curr_header->ME_flag = (flags >> 14) & 1;
  r->curr_header->data_length = get_be(b + 8, 8);
  memcpy(r->curr_header->type, b + 16, 128); r->curr_header->type[128] = 0;
  r->bytes_total = r->curr_header->data_length; r->bytes_left = r->bytes_total; r->bytes_pad = pad8(r->bytes_total);
  r->rec_start = (n_uint64_t)ftello(r->fp); r->rec_ptr = 0; r->header_nextP = 0;
  return LIME_SUCCESS;
}
char *limeReaderType(LimeReader *r) { return r->curr_header->type; }
n_uint64_t limeReaderBytes(LimeReader *r) { return r->bytes_total; }
int limeReaderReadData(void *dest, n_uint64_t *nbytes, LimeReader *r) {
  n_uint64_t n = *nbytes;
  int st = LIME_SUCCESS;
  if (n > r->bytes_left) { n = r->bytes_left; st = LIME_EOR; }
  if (fread(dest, 1, n, r->fp) != n) return LIME_ERR_READ;
  *nbytes = n; r->bytes_left -= n; r->rec_ptr += n;
  return st;
}
int limeReaderSeek(LimeReader *r, off_t offset, int whence) {
  n_uint64_t pos = whence == SEEK_SET ? (n_uint64_t)offset : (whence == SEEK_CUR ? r->rec_ptr + (n_uint64_t)offset : r->bytes_total + (n_uint64_t)offset);
  if (pos > r->bytes_total) return LIME_ERR_SEEK;
  if (fseeko(r->fp, (off_t)(r->rec_start + pos), SEEK_SET) != 0) return LIME_ERR_SEEK;
  r->rec_ptr = pos; r->bytes_left = r->bytes_total - pos;
  return LIME_SUCCESS;
}
