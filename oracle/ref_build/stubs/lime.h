/* Stand-in for the un-vendored third-party c-lime library (usqcd-software/c-lime, "tested with 1.2.3",
 * reference README:10), which is absent from /root/reference and from this image.  It declares - and
 * lime_standin.c implements - the subset of the c-lime API that the reference's io/ directory calls
 * (io/selector.h:37-57), following the published LIME record format:
 *   144-byte header = magic 0x456789AB (u32 BE), version 1 (u16 BE), flags (u16 BE: bit 15 MB, bit 14 ME),
 *   data length (u64 BE), type (128 bytes, NUL padded); data padded with zeros to a multiple of 8 bytes.
 * The CONTAINER layer is therefore a restatement on both sides of every I/O parity test ("parity unpinned"
 * for LIME itself); what those tests pin through the reference's own unmodified reader/writer code is
 * everything inside the records: ILDG payload order and endianness, the XML records, the SciDAC checksum.
 * TEST INFRASTRUCTURE ONLY (oracle/ref_build). */
#ifndef TMB_REF_STUB_LIME_H
#define TMB_REF_STUB_LIME_H
#include <stdint.h>
#include <stdio.h>
#include <sys/types.h>
typedef uint64_t n_uint64_t;
#define LIME_SUCCESS 0
#define LIME_ERR_LAST_NOT_WRITTEN (-1)
#define LIME_ERR_PARAM (-2)
#define LIME_ERR_HEADER_NEXT (-3)
#define LIME_LAST_REC_WRITTEN (-4)
#define LIME_ERR_WRITE (-5)
#define LIME_EOR (-6)
#define LIME_EOF (-7)
#define LIME_ERR_READ (-8)
#define LIME_ERR_SEEK (-9)
#define LIME_ERR_MBME (-10)
#define LIME_ERR_CLOSE (-11)
typedef struct { unsigned int lime_version; int MB_flag, ME_flag; char *type; n_uint64_t data_length; } LimeRecordHeader;
typedef struct {
  int first_record, last_written, header_nextP;
  FILE *fp;
  n_uint64_t bytes_total, bytes_left, rec_ptr, rec_start, bytes_pad;
  int isLastP;
} LimeWriter;
typedef struct {
  int first_read, is_last, header_nextP;
  FILE *fp;
  LimeRecordHeader *curr_header;
  n_uint64_t bytes_left, bytes_total, rec_ptr, rec_start, bytes_pad;
} LimeReader;
LimeRecordHeader *limeCreateHeader(int MB_flag, int ME_flag, char *type, n_uint64_t reclen);
void limeDestroyHeader(LimeRecordHeader *h);
LimeWriter *limeCreateWriter(FILE *fp);
int limeDestroyWriter(LimeWriter *w);
int limeWriteRecordHeader(LimeRecordHeader *props, LimeWriter *w);
int limeWriteRecordData(void *source, n_uint64_t *nbytes, LimeWriter *w);
int limeWriterCloseRecord(LimeWriter *w);
LimeReader *limeCreateReader(FILE *fp);
void limeDestroyReader(LimeReader *r);
int limeReaderNextRecord(LimeReader *r);
char *limeReaderType(LimeReader *r);
n_uint64_t limeReaderBytes(LimeReader *r);
int limeReaderReadData(void *dest, n_uint64_t *nbytes, LimeReader *r);
int limeReaderSeek(LimeReader *r, off_t offset, int whence);
int limeReaderCloseRecord(LimeReader *r);
#endif
