/*
 * hcj_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY, see hcj_oracle.h).
 *
 * Literal restatement of hardcamls/video-coding's OCaml JPEG model.  Each
 * function cites the reference file:line it follows.  Clarity over speed: the
 * bit reader peeks bit-by-bit exactly as the model does.
 */
#include "hcj_oracle.h"

#include <math.h>
#include <setjmp.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int64_t i64;

/* OCaml exceptions -> longjmp to the API entry point. */
static __thread jmp_buf *orc_jmp;
static void orc_raise(int code) { longjmp(*orc_jmp, code); }
#define ORC_TRY(status_var)          \
  jmp_buf jb_;                       \
  jmp_buf *saved_jb_ = orc_jmp;      \
  orc_jmp = &jb_;                    \
  int status_var = setjmp(jb_);      \
  if (status_var == 0)
#define ORC_END_TRY orc_jmp = saved_jb_

static i64 asr(i64 x, int n) { return x >> n; } /* arithmetic on every supported compiler; checked in tests */
static i64 round_up(i64 x, i64 m) { return ((x + m - 1) / m) * m; } /* Int.round_up, x >= 0 */

/* ------------------------------------------------------------------------------------------
 * Bitstream_reader (common/src/bitstream_reader.ml)
 * ---------------------------------------------------------------------------------------- */
void orc_bits_create(orc_bits *b, const uint8_t *buf, i64 len) { /* :16 */
  b->buf = buf;
  b->len = len;
  b->length_in_bits = len * 8;
  b->bit_pos = 0;
}

static int get_byte(const orc_bits *b, i64 byte_no) { /* :19-22: out of range reads as '\000' */
  if (byte_no < 0 || byte_no >= b->len) return 0;
  return b->buf[byte_no];
}

static int get_bit(const orc_bits *b, i64 pos) { /* :24-29 */
  i64 byte_no = pos >> 3;
  int bit_no = 7 - (int)(pos & 7);
  return (get_byte(b, byte_no) >> bit_no) & 1;
}

static i64 show(orc_bits *b, int n) { /* :31-38 */
  if (n >= b->length_in_bits) orc_raise(ORC_ERR_BITS_OUT_OF_BOUNDS);
  uint64_t v = 0;
  for (int i = 0; i < n; i++) v = (v << 1) | (uint64_t)get_bit(b, b->bit_pos + i);
  return (i64)v;
}
static void advance(orc_bits *b, i64 n) { b->bit_pos += n; } /* :40 */
static i64 get(orc_bits *b, int n) {                         /* :42-46 */
  i64 v = show(b, n);
  advance(b, n);
  return v;
}
static void align_to_byte(orc_bits *b) { /* :51-54 */
  int num_bits = (int)(b->bit_pos & 7);
  if (num_bits != 0) advance(b, 8 - num_bits);
}

int orc_bits_show(orc_bits *b, int n, i64 *v) {
  ORC_TRY(st) { *v = show(b, n); }
  ORC_END_TRY;
  return st;
}
int orc_bits_get(orc_bits *b, int n, i64 *v) {
  ORC_TRY(st) { *v = get(b, n); }
  ORC_END_TRY;
  return st;
}
void orc_bits_advance(orc_bits *b, i64 n) { advance(b, n); }
void orc_bits_align_to_byte(orc_bits *b) { align_to_byte(b); }

/* ------------------------------------------------------------------------------------------
 * Bitstream_writer (common/src/bitstream_writer.ml)
 * ---------------------------------------------------------------------------------------- */
void orc_writer_create(orc_writer *w) { /* :11-17 */
  w->word_buffer = 0;
  w->word_bits = 0;
  w->capacity = 16 * 1024;
  w->buffer = (uint8_t *)malloc((size_t)w->capacity);
  w->bytes_written = 0;
}
void orc_writer_free(orc_writer *w) {
  free(w->buffer);
  w->buffer = NULL;
}
static void writer_add_char(orc_writer *w, int c) {
  if (w->bytes_written >= w->capacity) {
    w->capacity *= 2;
    w->buffer = (uint8_t *)realloc(w->buffer, (size_t)w->capacity);
  }
  w->buffer[w->bytes_written++] = (uint8_t)c;
}
static void writer_flush(orc_writer *w, int stuffing) { /* :19-30 */
  while (w->word_bits >= 8) {
    int d = (int)((w->word_buffer >> (w->word_bits - 8)) & 0xff);
    writer_add_char(w, d);
    w->word_bits -= 8;
    if (stuffing && d == 0xff) writer_add_char(w, 0);
  }
}
void orc_writer_put_bits(orc_writer *w, int stuffing, i64 value, int bits) { /* :32-40 */
  if (bits > 16) orc_raise(ORC_ERR_INVALID_ARG); /* assert (bits <= 16) */
  if (bits == 0) return;
  w->word_buffer = (w->word_buffer << bits) | ((uint64_t)value & ((1ull << bits) - 1));
  w->word_bits += bits;
  writer_flush(w, stuffing);
}
static i64 bits_written(const orc_writer *w) { return w->bytes_written * 8 + w->word_bits; } /* :43 */
void orc_writer_flush_with_1s(orc_writer *w, int stuffing) {                                  /* :45-49 */
  /* NB (:43): bytes_written counts stuffed zeros too; only (bits_written land 7) matters. */
  while ((bits_written(w) & 7) != 0) orc_writer_put_bits(w, stuffing, 1, 1);
}

/* ------------------------------------------------------------------------------------------
 * Markers (jpeg/model/src/markers.ml) and Decoder.Header (decoder.ml:5-71)
 * ---------------------------------------------------------------------------------------- */
enum { M_SOF0 = 0xc0, M_DHT = 0xc4, M_SOI = 0xd8, M_EOI = 0xd9, M_SOS = 0xda, M_DQT = 0xdb, M_DRI = 0xdd,
       M_APP0 = 0xe0, M_APP15 = 0xef, M_COM = 0xfe, M_RST0 = 0xd0 };

static void component_decode(orc_bits *b, orc_component *c) { /* markers.ml:15-25 */
  c->identifier = (int)get(b, 8);
  c->horizontal_sampling_factor = (int)get(b, 4);
  c->vertical_sampling_factor = (int)get(b, 4);
  c->quantization_table_identifier = (int)get(b, 8);
}
static void sof_decode(orc_bits *b, orc_sof *s) { /* markers.ml:49-59 */
  s->present = 1;
  s->length = (int)get(b, 16);
  s->sample_precision = (int)get(b, 8);
  s->height = (int)get(b, 16);
  s->width = (int)get(b, 16);
  s->number_of_components = (int)get(b, 8);
  if (s->number_of_components > 4) orc_raise(ORC_ERR_UNSUPPORTED_GEOMETRY); /* stated domain limit */
  for (int i = 0; i < s->number_of_components; i++) component_decode(b, &s->components[i]);
}
static void sos_decode(orc_bits *b, orc_sos *s) { /* markers.ml:84-89,111-129 */
  s->present = 1;
  s->length = (int)get(b, 16);
  s->number_of_image_components = (int)get(b, 8);
  if (s->number_of_image_components > 4) orc_raise(ORC_ERR_UNSUPPORTED_GEOMETRY); /* stated domain limit */
  for (int i = 0; i < s->number_of_image_components; i++) {
    s->scan_components[i].selector = (int)get(b, 8);
    s->scan_components[i].dc_coef_selector = (int)get(b, 4);
    s->scan_components[i].ac_coef_selector = (int)get(b, 4);
  }
  s->start_of_predictor_selection = (int)get(b, 8);
  s->end_of_predictor_selection = (int)get(b, 8);
  s->successive_approximation_bit_high = (int)get(b, 4);
  s->successive_approximation_bit_low = (int)get(b, 4);
}
static void dqt_table_decode(orc_bits *b, orc_dqt *q);
static void dqt_decode(orc_bits *b, orc_dqt *q) { /* markers.ml:162-168: ONE table per segment */
  q->length = (int)get(b, 16);
  dqt_table_decode(b, q);
}
static void dqt_table_decode(orc_bits *b, orc_dqt *q) {
  int pq = (int)get(b, 4);
  if (pq > 1) orc_raise(ORC_ERR_UNSUPPORTED_GEOMETRY); /* stated domain limit: 8- or 16-bit elements only */
  q->element_precision = 8 << pq;
  q->table_identifier = (int)get(b, 4);
  for (int i = 0; i < 64; i++) q->elements[i] = get(b, q->element_precision);
}
static void dht_table_decode(orc_bits *b, orc_dht *h);
static void dht_decode(orc_bits *b, orc_dht *h) { /* markers.ml:210-220: ONE table per segment */
  h->length = (int)get(b, 16);
  dht_table_decode(b, h);
}
static void dht_table_decode(orc_bits *b, orc_dht *h) {
  h->table_class = (int)get(b, 4);
  h->destination_identifier = (int)get(b, 4);
  int total = 0;
  for (int i = 0; i < 16; i++) {
    h->lengths[i] = (int)get(b, 8);
    total += h->lengths[i];
  }
  if (total > 256) orc_raise(ORC_ERR_BAD_HUFFMAN_TABLE); /* stated domain limit: more symbols than byte values */
  h->nvalues = total;
  for (int i = 0; i < total; i++) h->values[i] = (int)get(b, 8);
}

static void find_marker(orc_bits *b) { /* decoder.ml:24-29 */
  align_to_byte(b);
  for (;;) {
    /* The model would spin forever on zero-extended input; report it instead. */
    if (b->bit_pos >= b->length_in_bits) orc_raise(ORC_ERR_TRUNCATED);
    if (get(b, 8) == 0xff) return;
  }
}

/* flags & ORC_FLAG_T81_TABLES (stated extension): a DQT / DHT segment holds as many tables as
 * its length field covers (T.81 B.2.4.1, B.2.4.2) and the cursor continues at the end of the segment; 0xFF fill
 * bytes in front of a marker code are skipped (T.81 B.1.1.2).  Without it: the model verbatim. */
static void header_decode(orc_bits *b, orc_header *h, int flags) { /* decoder.ml:37-70 */
  const int t81 = (flags & ORC_FLAG_T81_TABLES) != 0;
  memset(h, 0, sizeof(*h));
  for (;;) {
    find_marker(b);
    int code = (int)get(b, 8);
    while (t81 && code == 0xff && b->bit_pos < b->length_in_bits) code = (int)get(b, 8);

    if (code == M_SOF0) {
      sof_decode(b, &h->frame);
    } else if (code == M_SOS) {
      sos_decode(b, &h->scan);
      h->scan_bit_pos = b->bit_pos;
      return;
    } else if (code == M_DQT) { /* list cons: newest first (:51) */
      if (h->n_quant_tables == ORC_MAX_TABLES) orc_raise(ORC_ERR_UNSUPPORTED_GEOMETRY);
      memmove(&h->quant_tables[1], &h->quant_tables[0], sizeof(orc_dqt) * (size_t)h->n_quant_tables);
      const i64 seg_end = b->bit_pos + 8 * show(b, 16);
      dqt_decode(b, &h->quant_tables[0]);
      h->n_quant_tables++;
      while (t81 && b->bit_pos < seg_end) { /* further tables: Pq/Tq + 64 elements each, no length field */
        if (h->n_quant_tables == ORC_MAX_TABLES) orc_raise(ORC_ERR_UNSUPPORTED_GEOMETRY);
        memmove(&h->quant_tables[1], &h->quant_tables[0], sizeof(orc_dqt) * (size_t)h->n_quant_tables);
        h->n_quant_tables++;
        dqt_table_decode(b, &h->quant_tables[0]);
      }
      if (t81 && b->bit_pos < seg_end) b->bit_pos = seg_end;
    } else if (code == M_DHT) { /* :55 */
      if (h->n_huffman_tables == ORC_MAX_TABLES) orc_raise(ORC_ERR_UNSUPPORTED_GEOMETRY);
      memmove(&h->huffman_tables[1], &h->huffman_tables[0], sizeof(orc_dht) * (size_t)h->n_huffman_tables);
      const i64 seg_end = b->bit_pos + 8 * show(b, 16);
      dht_decode(b, &h->huffman_tables[0]);
      h->n_huffman_tables++;
      while (t81 && b->bit_pos < seg_end) {
        if (h->n_huffman_tables == ORC_MAX_TABLES) orc_raise(ORC_ERR_UNSUPPORTED_GEOMETRY);
        memmove(&h->huffman_tables[1], &h->huffman_tables[0], sizeof(orc_dht) * (size_t)h->n_huffman_tables);
        h->n_huffman_tables++;
        dht_table_decode(b, &h->huffman_tables[0]);
      }
      if (t81 && b->bit_pos < seg_end) b->bit_pos = seg_end;
    } else if (code == M_DRI) { /* markers.ml:193-197 */
      h->restart_interval_present = 1;
      h->restart_interval_length = (int)get(b, 16);
      h->restart_interval = (int)get(b, 16);
    } else if (code == M_SOI) {
      /* continue */
    } else if ((code >= M_APP0 && code <= M_APP15) || code == M_COM) { /* skip, :31-34 */
      i64 len = show(b, 16);
      advance(b, len * 8);
    } else {
      orc_raise(ORC_ERR_UNSUPPORTED_MARKER);
    }
  }
}

int orc_header_decode_ex(const uint8_t *jpeg, i64 len, int flags, orc_header *h) {
  orc_bits b;
  orc_bits_create(&b, jpeg, len);
  ORC_TRY(st) { header_decode(&b, h, flags); }
  ORC_END_TRY;
  return st;
}
int orc_header_decode(const uint8_t *jpeg, i64 len, orc_header *h) { return orc_header_decode_ex(jpeg, len, 0, h); }

/* decoder.ml:261-281.  Stops at the first FF xx with xx != 00.  With restart_ext, FF D0..D7 ends
 * an interval instead (stated extension); seg_off (if non-NULL) receives the destuffed offset at
 * which each interval starts, nseg their count. */
static int extract_entropy(const uint8_t *buf, i64 len, i64 pos, int restart_ext, uint8_t *out, i64 *out_len,
                           i64 *seg_off, i64 max_seg, i64 *nseg) {
  i64 n = 0, ns = 0;
  int prev = 0x00;
  if (seg_off && ns < max_seg) seg_off[ns] = 0;
  ns++;
  for (;; pos++) {
    if (pos >= len + 2) return ORC_ERR_NO_TERMINATOR; /* model: zero-extends and never stops */
    int c = (pos < len) ? buf[pos] : 0;
    if (prev == 0xff) {
      if (c == 0x00) {
        out[n++] = 0xff;
        prev = c;
      } else if (restart_ext && c >= M_RST0 && c <= M_RST0 + 7) {
        if (seg_off && ns < max_seg) seg_off[ns] = n;
        ns++;
        prev = 0x00;
      } else {
        break;
      }
    } else if (c == 0xff) {
      prev = c;
    } else {
      out[n++] = (uint8_t)c;
      prev = c;
    }
  }
  *out_len = n;
  if (nseg) *nseg = ns;
  return ORC_OK;
}

int orc_extract_entropy_coded_bits(const uint8_t *jpeg, i64 len, i64 start_byte, uint8_t *out, i64 *out_len) {
  return extract_entropy(jpeg, len, start_byte, 0, out, out_len, NULL, 0, NULL);
}

/* ------------------------------------------------------------------------------------------
 * Tables (jpeg/model/src/tables.ml)
 * ---------------------------------------------------------------------------------------- */
int orc_create_code_table(const int lengths[16], const int *values, orc_code *codes) { /* :27-45 */
  int n = 0, data_pos = 0;
  i64 code = 0;
  for (int length_pos = 0; length_pos < 16; length_pos++) {
    if (lengths[length_pos] == 0) {
      code = code << 1;
    } else {
      for (int i = 0; i < lengths[length_pos]; i++) {
        codes[n].length = length_pos + 1;
        codes[n].bits = (int)(code + i);
        codes[n].data = values[data_pos + i];
        n++;
      }
      code = (code + lengths[length_pos]) << 1;
      data_pos += lengths[length_pos];
    }
  }
  return n;
}

/* Default specs = ITU-T T.81 Annex K tables K.3-K.6 (tables.ml:54-476). */
static const int dc_luma_lengths[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const int dc_chroma_lengths[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const int dc_values[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const int ac_luma_lengths[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const int ac_luma_values[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const int ac_chroma_lengths[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const int ac_chroma_values[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

void orc_default_spec(int which, const int **lengths, const int **values, int *nvalues) {
  switch (which) {
    case 0: *lengths = dc_luma_lengths; *values = dc_values; *nvalues = 12; break;
    case 1: *lengths = dc_chroma_lengths; *values = dc_values; *nvalues = 12; break;
    case 2: *lengths = ac_luma_lengths; *values = ac_luma_values; *nvalues = 162; break;
    default: *lengths = ac_chroma_lengths; *values = ac_chroma_values; *nvalues = 162; break;
  }
}

/* Tables.Lut (tables.ml:478-502): direct table over max_bits; length 0 == None. */
typedef struct {
  int max_bits;
  uint8_t *length; /* [1 << max_bits] */
  uint8_t *data;
} orc_lut;

static int lut_create(orc_lut *l, const orc_code *codes, int n) {
  int max_bits = 0;
  for (int i = 0; i < n; i++)
    if (codes[i].length > max_bits) max_bits = codes[i].length;
  l->max_bits = max_bits;
  size_t size = (size_t)1 << max_bits;
  l->length = (uint8_t *)calloc(size, 1);
  l->data = (uint8_t *)calloc(size, 1);
  for (int i = 0; i < n; i++) {
    int null_bits = max_bits - codes[i].length;
    i64 first = (i64)codes[i].bits << null_bits;
    i64 count = (i64)1 << null_bits;
    for (i64 j = first; j < first + count; j++) {
      if (j < 0 || j >= (i64)size) return ORC_ERR_BAD_HUFFMAN_TABLE; /* OCaml: index out of bounds (over-subscribed DHT) */
      l->length[j] = (uint8_t)codes[i].length;
      l->data[j] = (uint8_t)codes[i].data;
    }
  }
  return ORC_OK;
}
static void lut_free(orc_lut *l) {
  free(l->length);
  free(l->data);
  l->length = l->data = NULL;
}

/* Tables.Encoder.dc_table (tables.ml:505-514): codes sorted by category. */
static int encoder_dc_table(const int *lengths, const int *values, orc_code *out) {
  orc_code codes[256];
  int n = orc_create_code_table(lengths, values, codes);
  /* stable insertion sort by data (List.sort is a stable merge sort) */
  for (int i = 0; i < n; i++) out[i] = codes[i];
  for (int i = 1; i < n; i++) {
    orc_code c = out[i];
    int j = i - 1;
    while (j >= 0 && out[j].data > c.data) {
      out[j + 1] = out[j];
      j--;
    }
    out[j + 1] = c;
  }
  return n;
}
/* Tables.Encoder.ac_table (tables.ml:516-544): rows grouped by run (in sorted order, NOT indexed by
 * run value), each row indexed by position; a zero-length dummy is prepended when the row lacks size 0. */
static int encoder_ac_table(const int *lengths, const int *values, orc_code *out /*[16][16]*/, int row_len[16],
                            int *nrows) {
  orc_code codes[16 * 255];
  int n = orc_create_code_table(lengths, values, codes);
  for (int i = 0; i < n; i++) codes[i].data = ((codes[i].data >> 4) & 0xf) << 4 | (codes[i].data & 0xf);
  for (int i = 1; i < n; i++) { /* sort by (run, size) */
    orc_code c = codes[i];
    int j = i - 1;
    while (j >= 0 && codes[j].data > c.data) {
      codes[j + 1] = codes[j];
      j--;
    }
    codes[j + 1] = c;
  }
  int rows = 0;
  for (int i = 0; i < 16; i++) row_len[i] = 0;
  for (int i = 0; i < n;) {
    int run = codes[i].data >> 4;
    if (rows >= 16) return ORC_ERR_ENCODER_PARAMS;
    int k = 0;
    if ((codes[i].data & 0xf) != 0) {
      orc_code d = {0, 0, 0};
      out[rows * 16 + k++] = d;
    }
    while (i < n && (codes[i].data >> 4) == run) {
      if (k >= 16) return ORC_ERR_ENCODER_PARAMS;
      out[rows * 16 + k++] = codes[i++];
    }
    row_len[rows++] = k;
  }
  *nrows = rows;
  return ORC_OK;
}
int orc_encoder_dc_table(int which, orc_code *out, int *n) {
  const int *l, *v;
  int nv;
  orc_default_spec(which, &l, &v, &nv);
  *n = encoder_dc_table(l, v, out);
  return ORC_OK;
}
int orc_encoder_ac_table(int which, orc_code *out, int row_len[16]) {
  const int *l, *v;
  int nv, nrows;
  orc_default_spec(2 + which, &l, &v, &nv); /* which: 0 = ac_luma, 1 = ac_chroma */
  return encoder_ac_table(l, v, out, row_len, &nrows);
}

/* ------------------------------------------------------------------------------------------
 * Zigzag (zigzag.ml) / Quant_tables (quant_tables.ml)
 * ---------------------------------------------------------------------------------------- */
const int orc_zigzag_inverse[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                    12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const int orc_zigzag_forward[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42,
                                    3,  8,  12, 17, 25, 30, 41, 43, 9,  11, 18, 24, 31, 40, 44, 53,
                                    10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60,
                                    21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

/* Annex K tables held in natural order but USED as if zig-zag (quant_tables.ml:3-137). */
static const int quant_luma[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                   14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                   18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                   49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const int quant_chroma[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                                     24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

void orc_quant_scale(int chroma, int quality, i64 out[64]) { /* quant_tables.ml:139-147 */
  const int *table = chroma ? quant_chroma : quant_luma;
  i64 q = quality < 1 ? 1 : quality > 100 ? 100 : quality;
  i64 s = q < 50 ? 5000 / q : 200 - 2 * q;
  for (int i = 0; i < 64; i++) {
    i64 d = (table[i] * s + 50) / 100;
    out[i] = d < 1 ? 1 : d > 255 ? 255 : d;
  }
}

/* ------------------------------------------------------------------------------------------
 * Dct.Chen (dct.ml:3-197)
 * ---------------------------------------------------------------------------------------- */
enum { W1 = 2841, W2 = 2676, W3 = 2408, W5 = 1609, W6 = 1108, W7 = 565 };

/* OCaml lsl on a negative int == multiply by 2^n (two's complement). */
static i64 lsl(i64 x, int n) { return (i64)((uint64_t)x << n); }

static void idct_row(i64 *blk) { /* dct.ml:11-54 */
  i64 x0 = lsl(blk[0], 11) + 128, x1 = lsl(blk[4], 11), x2 = blk[6], x3 = blk[2], x4 = blk[1], x5 = blk[7],
      x6 = blk[5], x7 = blk[3], x8;
  x8 = W7 * (x4 + x5);
  x4 = x8 + (W1 - W7) * x4;
  x5 = x8 - (W1 + W7) * x5;
  x8 = W3 * (x6 + x7);
  x6 = x8 - (W3 - W5) * x6;
  x7 = x8 - (W3 + W5) * x7;
  x8 = x0 + x1;
  x0 = x0 - x1;
  x1 = W6 * (x3 + x2);
  x2 = x1 - (W2 + W6) * x2;
  x3 = x1 + (W2 - W6) * x3;
  x1 = x4 + x6;
  x4 = x4 - x6;
  x6 = x5 + x7;
  x5 = x5 - x7;
  x7 = x8 + x3;
  x8 = x8 - x3;
  x3 = x0 + x2;
  x0 = x0 - x2;
  x2 = asr(181 * (x4 + x5) + 128, 8);
  x4 = asr(181 * (x4 - x5) + 128, 8);
  blk[0] = asr(x7 + x1, 8);
  blk[1] = asr(x3 + x2, 8);
  blk[2] = asr(x0 + x4, 8);
  blk[3] = asr(x8 + x6, 8);
  blk[4] = asr(x8 - x6, 8);
  blk[5] = asr(x0 - x4, 8);
  blk[6] = asr(x3 - x2, 8);
  blk[7] = asr(x7 - x1, 8);
}

static void idct_col(i64 *blk) { /* dct.ml:56-98 (blk points at the column head, stride 8) */
  i64 x0 = lsl(blk[8 * 0], 8) + 8192, x1 = lsl(blk[8 * 4], 8), x2 = blk[8 * 6], x3 = blk[8 * 2], x4 = blk[8 * 1],
      x5 = blk[8 * 7], x6 = blk[8 * 5], x7 = blk[8 * 3], x8;
  x8 = W7 * (x4 + x5) + 4;
  x4 = asr(x8 + (W1 - W7) * x4, 3);
  x5 = asr(x8 - (W1 + W7) * x5, 3);
  x8 = W3 * (x6 + x7) + 4;
  x6 = asr(x8 - (W3 - W5) * x6, 3);
  x7 = asr(x8 - (W3 + W5) * x7, 3);
  x8 = x0 + x1;
  x0 = x0 - x1;
  x1 = W6 * (x3 + x2) + 4;
  x2 = asr(x1 - (W2 + W6) * x2, 3);
  x3 = asr(x1 + (W2 - W6) * x3, 3);
  x1 = x4 + x6;
  x4 = x4 - x6;
  x6 = x5 + x7;
  x5 = x5 - x7;
  x7 = x8 + x3;
  x8 = x8 - x3;
  x3 = x0 + x2;
  x0 = x0 - x2;
  x2 = asr(181 * (x4 + x5) + 128, 8);
  x4 = asr(181 * (x4 - x5) + 128, 8);
  blk[8 * 0] = asr(x7 + x1, 14);
  blk[8 * 1] = asr(x3 + x2, 14);
  blk[8 * 2] = asr(x0 + x4, 14);
  blk[8 * 3] = asr(x8 + x6, 14);
  blk[8 * 4] = asr(x8 - x6, 14);
  blk[8 * 5] = asr(x0 - x4, 14);
  blk[8 * 6] = asr(x3 - x2, 14);
  blk[8 * 7] = asr(x7 - x1, 14);
}

void orc_chen_inverse_8x8(i64 block[64]) { /* dct.ml:100-107: all rows, then all columns */
  for (int i = 0; i < 8; i++) idct_row(block + 8 * i);
  for (int i = 0; i < 8; i++) idct_col(block + i);
}

static i64 c4(i64 f, i64 g) { return asr(362 * (f + g), 9); }        /* dct.ml:109 */
static i64 c62(i64 f, i64 g) { return asr(196 * f + 473 * g, 9); }   /* :110 */
static i64 c71(i64 f, i64 g) { return asr(100 * f + 502 * g, 9); }   /* :111 */
static i64 c35(i64 f, i64 g) { return asr(426 * f + 284 * g, 9); }   /* :112 */

static void dct_1d(i64 *b, int s) { /* dct.ml:114-149 (s = 8, columns) and :151-187 (s = 1, rows) */
  i64 a0 = b[0 * s] + b[7 * s], c3 = b[0 * s] - b[7 * s];
  i64 a1 = b[1 * s] + b[6 * s], c2 = b[1 * s] - b[6 * s];
  i64 a2 = b[2 * s] + b[5 * s], c1 = b[2 * s] - b[5 * s];
  i64 a3 = b[3 * s] + b[4 * s], c0 = b[3 * s] - b[4 * s];
  i64 b0 = a0 + a3, b1 = a1 + a2, b2 = a1 - a2, b3 = a0 - a3;
  b[0 * s] = c4(b0, b1);
  b[4 * s] = c4(b0, -b1);
  b[2 * s] = c62(b2, b3);
  b[6 * s] = c62(b3, -b2);
  b0 = c4(c2, -c1);
  b1 = c4(c2, c1);
  a0 = c0 + b0;
  a1 = c0 - b0;
  a2 = c3 - b1;
  a3 = c3 + b1;
  b[1 * s] = c71(a0, a3);
  b[5 * s] = c35(a1, a2);
  b[3 * s] = c35(a2, -a1);
  b[7 * s] = c71(a3, -a0);
}

void orc_chen_forward_8x8(i64 block[64]) { /* dct.ml:189-196: all columns, then all rows */
  for (int i = 0; i < 8; i++) dct_1d(block + i, 8);
  for (int i = 0; i < 8; i++) dct_1d(block + 8 * i, 1);
}

/* ------------------------------------------------------------------------------------------
 * Codewords: Encoder.size / magnitude (encoder.ml:143-147), Decoder.mag' (decoder.ml:73-79)
 * ---------------------------------------------------------------------------------------- */
int orc_size(i64 value) {
  if (value == 0) return 0;
  uint64_t a = (uint64_t)(value < 0 ? -value : value);
  int l = 0;
  while (a >>= 1) l++; /* Int.floor_log2 */
  return l + 1;
}
i64 orc_magnitude(int size, i64 value) {
  i64 mask = ((i64)1 << size) - 1;
  return value >= 0 ? (value & mask) : ((value - 1) & mask);
}
i64 orc_mag(int cat, i64 code) {
  if (cat > 0 && (code & ((i64)1 << (cat - 1)))) return code;
  return (code | lsl(-1, cat)) + 1;
}

int orc_rle(const i64 quant[64], i64 *dc_pred, i64 runs[65], i64 values[65]) { /* encoder.ml:127-141 */
  int n = 0;
  runs[n] = 0;
  values[n] = quant[0] - *dc_pred;
  n++;
  i64 run = 0;
  for (int pos = 1; pos < 64; pos++) {
    i64 value = quant[pos];
    if (pos == 63) {
      runs[n] = run;
      values[n] = value;
      n++;
    } else if (value != 0) {
      runs[n] = run;
      values[n] = value;
      n++;
      run = 0;
    } else {
      run++;
    }
  }
  *dc_pred = quant[0];
  return n;
}

/* ------------------------------------------------------------------------------------------
 * Decoder (decoder.ml)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int hs, vs;
  int decoded_width, decoded_height, actual_width, actual_height;
  i64 dc_pred;
  const i64 *quant_table;
  orc_lut dc_tab, ac_tab;
  uint8_t *plane;
} dec_component;

static i64 mag(orc_bits *b, int cat) { /* decoder.ml:81-87 */
  if (cat == 0) return 0;
  i64 v = get(b, cat);
  return orc_mag(cat, v);
}

static void huffman_decode(orc_bits *b, i64 *coefs, const orc_lut *dc_tab, const orc_lut *ac_tab) { /* :118-140 */
  /* dc_code (:89-96) */
  i64 code = show(b, dc_tab->max_bits);
  if (dc_tab->length[code] == 0) orc_raise(ORC_ERR_NO_DC_CODE);
  advance(b, dc_tab->length[code]);
  coefs[0] = mag(b, dc_tab->data[code]);
  int cof_cnt = 1;
  while (cof_cnt < 64) {
    /* ac_code (:98-105) */
    code = show(b, ac_tab->max_bits);
    if (ac_tab->length[code] == 0) orc_raise(ORC_ERR_NO_AC_CODE);
    advance(b, ac_tab->length[code]);
    int run = (ac_tab->data[code] >> 4) & 0xf, size = ac_tab->data[code] & 0xf;
    i64 m = mag(b, size);
    if (m == 0 && run == 0) {
      cof_cnt = 64;
    } else {
      cof_cnt += run;
      if (cof_cnt >= 64) orc_raise(ORC_ERR_COEF_INDEX);
      coefs[cof_cnt] = m;
      cof_cnt++;
    }
  }
}

static const orc_dht *find_huffman_table(const orc_header *h, int ac_dc, int id) { /* :238-245 */
  for (int i = 0; i < h->n_huffman_tables; i++)
    if (h->huffman_tables[i].table_class == ac_dc && h->huffman_tables[i].destination_identifier == id)
      return &h->huffman_tables[i];
  orc_raise(ORC_ERR_NO_HUFFMAN_TABLE);
  return NULL;
}

static void build_lut(const orc_header *h, int ac_dc, int id, orc_lut *lut) { /* :247-259 */
  const orc_dht *t = find_huffman_table(h, ac_dc, id);
  orc_code *codes = (orc_code *)malloc(sizeof(orc_code) * (size_t)(t->nvalues + 1));
  int n = orc_create_code_table(t->lengths, t->values, codes);
  if (ac_dc == 0) { /* stated domain limit: DC categories above 15 are not supported */
    for (int i = 0; i < n; i++)
      if (codes[i].data > 15) {
        free(codes);
        orc_raise(ORC_ERR_UNSUPPORTED_GEOMETRY);
      }
  }
  int st = lut_create(lut, codes, n);
  free(codes);
  if (st != ORC_OK) orc_raise(st);
}

static int decode_impl(const uint8_t *jpeg, i64 len, int flags, int want_blocks, orc_decoded *out,
                       dec_component *comps, uint8_t **entropy_buf, i64 **seg_off_buf) {
  orc_header *h = (orc_header *)malloc(sizeof(orc_header));
  orc_bits bits;
  orc_bits_create(&bits, jpeg, len);
  ORC_TRY(st) {
    header_decode(&bits, h, flags);
    /* init (:304-345) */
    if (!h->frame.present || !h->scan.present) orc_raise(ORC_ERR_NO_FRAME_OR_SCAN);
    const orc_sof *frame = &h->frame;
    const orc_sos *scan = &h->scan;
    int max_h = 0, max_v = 0; /* :294-302 */
    for (int i = 0; i < frame->number_of_components; i++) {
      if (frame->components[i].horizontal_sampling_factor > max_h) max_h = frame->components[i].horizontal_sampling_factor;
      if (frame->components[i].vertical_sampling_factor > max_v) max_v = frame->components[i].vertical_sampling_factor;
    }
    int ncomp = scan->number_of_image_components;
    if (ncomp < 1 || ncomp > 4 || max_h < 1 || max_v < 1 || max_h > 4 || max_v > 4)
      orc_raise(ORC_ERR_UNSUPPORTED_GEOMETRY);
    i64 rounded_width = round_up(frame->width, (i64)max_h * 8);
    i64 rounded_height = round_up(frame->height, (i64)max_v * 8);
    out->ncomp = ncomp;
    out->width = frame->width;
    out->height = frame->height;
    int bpm = 0;
    for (int i = 0; i < ncomp; i++) {
      const orc_component *c = NULL; /* find_component (:226-230) */
      for (int j = 0; j < frame->number_of_components; j++)
        if (frame->components[j].identifier == scan->scan_components[i].selector) {
          c = &frame->components[j];
          break;
        }
      if (!c) orc_raise(ORC_ERR_NO_COMPONENT);
      dec_component *d = &comps[i];
      d->hs = c->horizontal_sampling_factor;
      d->vs = c->vertical_sampling_factor;
      if (d->hs < 1 || d->vs < 1) orc_raise(ORC_ERR_UNSUPPORTED_GEOMETRY);
      d->decoded_width = (int)(rounded_width * d->hs / max_h);
      d->decoded_height = (int)(rounded_height * d->vs / max_v);
      d->actual_width = (int)((i64)frame->width * d->hs / max_h);
      d->actual_height = (int)((i64)frame->height * d->vs / max_v);
      d->dc_pred = 0;
      d->plane = (uint8_t *)calloc((size_t)d->decoded_width * (size_t)d->decoded_height + 1, 1); /* Plane.create */
      d->quant_table = NULL; /* find_quant_table (:232-236) */
      for (int j = 0; j < h->n_quant_tables; j++)
        if (h->quant_tables[j].table_identifier == c->quantization_table_identifier) {
          d->quant_table = h->quant_tables[j].elements;
          break;
        }
      if (!d->quant_table) orc_raise(ORC_ERR_NO_QUANT_TABLE);
      build_lut(h, 0, scan->scan_components[i].dc_coef_selector, &d->dc_tab);
      build_lut(h, 1, scan->scan_components[i].ac_coef_selector, &d->ac_tab);
      bpm += d->hs * d->vs;
      out->hs[i] = d->hs;
      out->vs[i] = d->vs;
      out->decoded_width[i] = d->decoded_width;
      out->decoded_height[i] = d->decoded_height;
      out->actual_width[i] = d->actual_width;
      out->actual_height[i] = d->actual_height;
    }
    if (bpm > 10) orc_raise(ORC_ERR_UNSUPPORTED_GEOMETRY);
    /* decode_seq geometry (:377-383) */
    int mcus_wide = comps[0].decoded_width / (8 * comps[0].hs);
    int mcus_high = comps[0].decoded_height / (8 * comps[0].vs);
    i64 nmcu = (i64)mcus_wide * mcus_high;
    i64 nblocks = nmcu * bpm;
    out->mcus_wide = mcus_wide;
    out->mcus_high = mcus_high;
    out->blocks_per_mcu = bpm;
    out->nblocks = nblocks;
    if (want_blocks) {
      out->coefs = (int32_t *)calloc((size_t)nblocks * 64 + 1, 4);
      out->dequant = (int32_t *)calloc((size_t)nblocks * 64 + 1, 4);
      out->recon = (uint8_t *)calloc((size_t)nblocks * 64 + 1, 1);
      out->dc_abs = (int32_t *)calloc((size_t)nblocks + 1, 4);
      out->block_comp = (int8_t *)calloc((size_t)nblocks + 1, 1);
    }

    /* extract_entropy_coded_bits (:261-281), per restart interval under the stated extension */
    int ri = (flags & ORC_FLAG_RESTART_EXT) && h->restart_interval_present ? h->restart_interval : 0;
    i64 expected_seg = ri > 0 ? (nmcu + ri - 1) / ri : 1;
    uint8_t *entropy = (uint8_t *)malloc((size_t)len + 16);
    *entropy_buf = entropy;
    i64 *seg_off = (i64 *)malloc(sizeof(i64) * (size_t)(expected_seg + 2));
    *seg_off_buf = seg_off;
    i64 entropy_len = 0, nseg = 0;
    int est = extract_entropy(jpeg, len, bits.bit_pos >> 3, ri > 0, entropy, &entropy_len, seg_off, expected_seg + 1, &nseg);
    if (est != ORC_OK) orc_raise(est);
    if (ri > 0 && nseg != expected_seg) orc_raise(ORC_ERR_RESTART_COUNT);
    seg_off[nseg <= expected_seg ? nseg : expected_seg + 1] = entropy_len;
    out->entropy_len = entropy_len;

    /* decode (:347-397): MCU raster -> scan components -> v x h blocks */
    i64 coefs[64], dequant[64], idct[64];
    orc_bits ebits;
    orc_bits_create(&ebits, entropy, ri > 0 ? seg_off[1] : entropy_len);
    i64 blk = 0, mcu = 0, seg = 0;
    for (int blk_y = 0; blk_y < mcus_high; blk_y++) {
      for (int blk_x = 0; blk_x < mcus_wide; blk_x++, mcu++) {
        if (ri > 0 && mcu > 0 && mcu % ri == 0) { /* extension: next interval, reset predictors */
          seg++;
          orc_bits_create(&ebits, entropy + seg_off[seg], seg_off[seg + 1] - seg_off[seg]);
          for (int i = 0; i < ncomp; i++) comps[i].dc_pred = 0;
        }
        for (int id = 0; id < ncomp; id++) {
          dec_component *c = &comps[id];
          for (int y = 0; y < c->vs; y++) {
            for (int x = 0; x < c->hs; x++, blk++) {
              int px = ((blk_x * c->hs) + x) * 8, py = ((blk_y * c->vs) + y) * 8; /* :367-368 */
              /* decode_coefficient_block (:151-165) */
              for (int i = 0; i < 64; i++) coefs[i] = 0;
              huffman_decode(&ebits, coefs, &c->dc_tab, &c->ac_tab);
              /* dequantize_dc_pred_and_inverse_zigzag (:142-149) */
              i64 dc = coefs[0] + c->dc_pred;
              dequant[0] = dc * c->quant_table[0];
              for (int i = 1; i < 64; i++) dequant[orc_zigzag_inverse[i]] = coefs[i] * c->quant_table[i];
              c->dc_pred = dc;
              memcpy(idct, dequant, sizeof(idct)); /* :357 */
              orc_chen_inverse_8x8(idct);          /* :358 */
              /* recon (:213-224) */
              for (int j = 0; j < 8; j++)
                for (int i = 0; i < 8; i++) {
                  int k = i + j * 8;
                  i64 v = idct[k] < -128 ? -128 : idct[k] > 127 ? 127 : idct[k];
                  if (px + i >= c->decoded_width || py + j >= c->decoded_height) orc_raise(ORC_ERR_PLANE_BOUNDS);
                  c->plane[(size_t)(px + i) + (size_t)(py + j) * (size_t)c->decoded_width] = (uint8_t)(v + 128);
                  if (want_blocks) out->recon[blk * 64 + k] = (uint8_t)(v + 128);
                }
              if (want_blocks) {
                for (int i = 0; i < 64; i++) {
                  out->coefs[blk * 64 + i] = (int32_t)coefs[i];
                  out->dequant[blk * 64 + i] = (int32_t)dequant[i];
                }
                out->dc_abs[blk] = (int32_t)dc;
                out->block_comp[blk] = (int8_t)id;
              }
            }
          }
        }
      }
    }
    /* get_decoded_planes (:399-401) and crop (:403-413) */
    for (int i = 0; i < ncomp; i++) {
      dec_component *c = &comps[i];
      out->plane[i] = c->plane;
      c->plane = NULL;
      out->cropped[i] = (uint8_t *)calloc((size_t)c->actual_width * (size_t)c->actual_height + 1, 1);
      int w = c->actual_width < c->decoded_width ? c->actual_width : c->decoded_width;
      int hh = c->actual_height < c->decoded_height ? c->actual_height : c->decoded_height;
      for (int row = 0; row < hh; row++) /* Plane.blit_available (plane.ml:22-34) */
        memcpy(out->cropped[i] + (size_t)row * (size_t)c->actual_width,
               out->plane[i] + (size_t)row * (size_t)c->decoded_width, (size_t)w);
    }
    /* get_yuv_frame (:415-420) -> Frame.of_planes -> infer_chroma_subsampling (frame.ml:42-61) */
    if (ncomp < 3) {
      out->yuv_status = ORC_ERR_NEED_3_COMPONENTS;
    } else {
      int yw = out->actual_width[0], yh = out->actual_height[0];
      int uw = out->actual_width[1], uh = out->actual_height[1];
      if (uw != out->actual_width[2] || uh != out->actual_height[2]) out->yuv_status = ORC_ERR_FRAME_INFER;
      else if (yw / 2 == uw && yh / 2 == uh) out->chroma = 420;
      else if (yw / 2 == uw && yh == uh) out->chroma = 422;
      else if (yw == uw && yh == uh) out->chroma = 444;
      else out->yuv_status = ORC_ERR_FRAME_INFER;
    }
  }
  ORC_END_TRY;
  free(h);
  return st;
}

int orc_decode(const uint8_t *jpeg, i64 len, int flags, int want_blocks, orc_decoded *out) {
  dec_component comps[4];
  memset(comps, 0, sizeof(comps));
  memset(out, 0, sizeof(*out));
  uint8_t *entropy = NULL;
  i64 *seg_off = NULL;
  int st = decode_impl(jpeg, len, flags, want_blocks, out, comps, &entropy, &seg_off);
  for (int i = 0; i < 4; i++) {
    free(comps[i].plane);
    lut_free(&comps[i].dc_tab);
    lut_free(&comps[i].ac_tab);
  }
  free(entropy);
  free(seg_off);
  out->status = st;
  if (st != ORC_OK) {
    int keep = st;
    orc_decoded_free(out);
    out->status = keep;
  }
  return st;
}

void orc_decoded_free(orc_decoded *d) {
  for (int i = 0; i < 4; i++) {
    free(d->plane[i]);
    free(d->cropped[i]);
    d->plane[i] = d->cropped[i] = NULL;
  }
  free(d->coefs);
  free(d->dc_abs);
  free(d->dequant);
  free(d->recon);
  free(d->block_comp);
  d->coefs = d->dc_abs = d->dequant = NULL;
  d->recon = NULL;
  d->block_comp = NULL;
}

/* ------------------------------------------------------------------------------------------
 * Encoder (encoder.ml)
 * ---------------------------------------------------------------------------------------- */
static void write_marker_code(orc_writer *w, int code) { /* encoder.ml:207-210 */
  orc_writer_put_bits(w, 0, 0xff, 8);
  orc_writer_put_bits(w, 0, code, 8);
}

typedef struct {
  int quant_table, dc_huffman_table, ac_huffman_table, component, hs, vs;
} enc_scan_component;

typedef struct { /* Encoder.Parameters.t (encoder.ml:287-369) */
  int width, height;
  int n_quant;
  i64 quant_tables[2][64];
  int n_huff; /* dc and ac tables come in pairs with identifiers 0.. */
  int ncomp;
  enc_scan_component sc[3];
} enc_params;

static void make_params(enc_params *p, int width, int height, int chroma, int quality) {
  static const int s420[6] = {2, 2, 1, 1, 1, 1}, s422[6] = {2, 2, 1, 2, 1, 2}, s444[6] = {1, 1, 1, 1, 1, 1};
  memset(p, 0, sizeof(*p));
  p->width = width;
  p->height = height;
  orc_quant_scale(0, quality, p->quant_tables[0]);
  if (chroma == 400) { /* monochrome (:351-368) */
    p->n_quant = 1;
    p->n_huff = 1;
    p->ncomp = 1;
    enc_scan_component c = {0, 0, 0, 1, 1, 1};
    p->sc[0] = c;
    return;
  }
  const int *s = chroma == 420 ? s420 : chroma == 422 ? s422 : chroma == 444 ? s444 : NULL;
  if (!s) orc_raise(ORC_ERR_ENCODER_PARAMS);
  orc_quant_scale(1, quality, p->quant_tables[1]);
  p->n_quant = 2;
  p->n_huff = 2;
  p->ncomp = 3;
  for (int i = 0; i < 3; i++) { /* :321-343 */
    enc_scan_component c = {i ? 1 : 0, i ? 1 : 0, i ? 1 : 0, i + 1, s[2 * i], s[2 * i + 1]};
    p->sc[i] = c;
  }
}

static void write_headers(const enc_params *p, orc_writer *w, int restart_interval) { /* encoder.ml:371-418 */
  write_marker_code(w, M_SOI);
  static const char tag[] = "Hardcaml JPEG."; /* write_app0 (:231-237) */
  write_marker_code(w, M_APP0);
  orc_writer_put_bits(w, 0, 2 + (int)strlen(tag), 16);
  for (size_t i = 0; i < strlen(tag); i++) orc_writer_put_bits(w, 0, tag[i], 8);
  for (int t = 0; t < p->n_quant; t++) { /* write_dqt (:224-229) + Dqt.encode (markers.ml:170-183) */
    write_marker_code(w, M_DQT);
    orc_writer_put_bits(w, 0, 3 + 64, 16);
    orc_writer_put_bits(w, 0, 0, 4);
    orc_writer_put_bits(w, 0, t, 4);
    for (int i = 0; i < 64; i++) orc_writer_put_bits(w, 0, p->quant_tables[t][i], 8);
  }
  write_marker_code(w, M_SOF0); /* write_sof (:239-250) + Sof.encode (markers.ml:61-71) */
  orc_writer_put_bits(w, 0, 2 + 6 + p->ncomp * 3, 16);
  orc_writer_put_bits(w, 0, 8, 8);
  orc_writer_put_bits(w, 0, p->height, 16);
  orc_writer_put_bits(w, 0, p->width, 16);
  orc_writer_put_bits(w, 0, p->ncomp, 8);
  for (int i = 0; i < p->ncomp; i++) {
    orc_writer_put_bits(w, 0, p->sc[i].component, 8);
    orc_writer_put_bits(w, 0, p->sc[i].hs, 4);
    orc_writer_put_bits(w, 0, p->sc[i].vs, 4);
    orc_writer_put_bits(w, 0, p->sc[i].quant_table, 8);
  }
  for (int cls = 0; cls < 2; cls++) /* all dc tables, then all ac tables (:405-408) */
    for (int t = 0; t < p->n_huff; t++) {
      const int *lengths, *values;
      int nvalues;
      orc_default_spec(cls * 2 + t, &lengths, &values, &nvalues);
      write_marker_code(w, M_DHT); /* write_dht (:212-222) + Dht.encode (markers.ml:222-231) */
      orc_writer_put_bits(w, 0, 3 + 16 + nvalues, 16);
      orc_writer_put_bits(w, 0, cls, 4);
      orc_writer_put_bits(w, 0, t, 4);
      for (int i = 0; i < 16; i++) orc_writer_put_bits(w, 0, lengths[i], 8);
      for (int i = 0; i < nvalues; i++) orc_writer_put_bits(w, 0, values[i], 8);
    }
  if (restart_interval > 0) { /* stated extension: DRI before SOS */
    write_marker_code(w, M_DRI);
    orc_writer_put_bits(w, 0, 4, 16);
    orc_writer_put_bits(w, 0, restart_interval, 16);
  }
  write_marker_code(w, M_SOS); /* write_sos (:252-264) + Sos.encode (markers.ml:131-150) */
  orc_writer_put_bits(w, 0, 2 + 4 + p->ncomp * 2, 16);
  orc_writer_put_bits(w, 0, p->ncomp, 8);
  for (int i = 0; i < p->ncomp; i++) {
    orc_writer_put_bits(w, 0, p->sc[i].component, 8);
    orc_writer_put_bits(w, 0, p->sc[i].dc_huffman_table, 4);
    orc_writer_put_bits(w, 0, p->sc[i].ac_huffman_table, 4);
  }
  orc_writer_put_bits(w, 0, 0, 8);
  orc_writer_put_bits(w, 0, 63, 8);
  orc_writer_put_bits(w, 0, 0, 4);
  orc_writer_put_bits(w, 0, 0, 4);
}

int orc_write_headers(int width, int height, int chroma, int quality, int restart_interval, uint8_t *out,
                      i64 cap, i64 *len) {
  orc_writer w;
  orc_writer_create(&w);
  enc_params p;
  ORC_TRY(st) {
    make_params(&p, width, height, chroma, quality);
    write_headers(&p, &w, restart_interval);
    if (w.bytes_written > cap) orc_raise(ORC_ERR_BUFFER_TOO_SMALL);
    memcpy(out, w.buffer, (size_t)w.bytes_written);
    *len = w.bytes_written;
  }
  ORC_END_TRY;
  orc_writer_free(&w);
  return st;
}

typedef struct {
  int hs, vs, pw, ph;
  uint8_t *plane;
  const i64 *quant_table;
  orc_code dc_table[256];
  int n_dc;
  orc_code ac_table[16 * 16];
  int ac_row_len[16], ac_rows;
  i64 dc_pred;
} enc_scan;

static i64 quant_and_scale(i64 fdct, i64 qnt) { /* encoder.ml:98-101: OCaml '/' truncates toward zero */
  return fdct < 0 ? (fdct - qnt * 2) / (qnt * 4) : (fdct + qnt * 2) / (qnt * 4);
}

static void write_ac(orc_writer *w, const enc_scan *s, int run, i64 value) { /* encoder.ml:162-168 */
  int size = orc_size(value);
  if (run >= s->ac_rows || size >= s->ac_row_len[run]) orc_raise(ORC_ERR_INVALID_ARG); /* OCaml index out of bounds */
  const orc_code *code = &s->ac_table[run * 16 + size];
  orc_writer_put_bits(w, 1, code->bits, code->length);
  orc_writer_put_bits(w, 1, orc_magnitude(size, value), size);
}

static int encode_impl(const uint8_t *const src[3], int width, int height, int chroma, int quality,
                       int restart_interval, int want_blocks, orc_encoded *out, orc_writer *w, enc_scan *scans) {
  ORC_TRY(st) {
    enc_params p;
    make_params(&p, width, height, chroma, quality);
    /* create (:437-472) */
    int max_h = 0, max_v = 0;
    for (int i = 0; i < p.ncomp; i++) {
      if (p.sc[i].hs > max_h) max_h = p.sc[i].hs;
      if (p.sc[i].vs > max_v) max_v = p.sc[i].vs;
    }
    int bpm = 0;
    for (int i = 0; i < p.ncomp; i++) {
      enc_scan *s = &scans[i];
      s->hs = p.sc[i].hs;
      s->vs = p.sc[i].vs;
      i64 pw = (i64)p.width * s->hs / max_h, ph = (i64)p.height * s->vs / max_v;
      s->pw = (int)round_up(pw, 8 * s->hs);
      s->ph = (int)round_up(ph, 8 * s->vs);
      s->plane = (uint8_t *)calloc((size_t)s->pw * (size_t)s->ph + 1, 1);
      s->quant_table = p.quant_tables[p.sc[i].quant_table];
      const int *l, *v;
      int nv;
      orc_default_spec(p.sc[i].dc_huffman_table, &l, &v, &nv);
      s->n_dc = encoder_dc_table(l, v, s->dc_table);
      orc_default_spec(2 + p.sc[i].ac_huffman_table, &l, &v, &nv);
      int r = encoder_ac_table(l, v, s->ac_table, s->ac_row_len, &s->ac_rows);
      if (r != ORC_OK) orc_raise(r);
      s->dc_pred = 0;
      bpm += s->hs * s->vs;
      /* Plane.blit_available src -> padded plane (:514-516); source plane dims per Frame.create (frame.ml:32-40) */
      int sw = i == 0 ? width : (chroma == 444 ? width : width / 2);
      int sh = i == 0 ? height : (chroma == 420 ? height / 2 : height);
      int cw = sw < s->pw ? sw : s->pw, ch = sh < s->ph ? sh : s->ph;
      if (chroma == 400) { /* encode_monochrome uses Plane.blit: ONE linear copy (encoder.ml:548, plane.ml:20) */
        memcpy(s->plane, src[i], (size_t)sw * (size_t)sh);
      } else {
        for (int row = 0; row < ch; row++)
          memcpy(s->plane + (size_t)row * (size_t)s->pw, src[i] + (size_t)row * (size_t)sw, (size_t)cw);
      }
    }
    write_headers(&p, w, restart_interval);
    /* encode_seq (:476-505) */
    int mbs_wide = scans[0].pw / (8 * scans[0].hs), mbs_high = scans[0].ph / (8 * scans[0].vs);
    i64 nblocks = (i64)mbs_wide * mbs_high * bpm;
    out->nblocks = nblocks;
    if (want_blocks) {
      out->quant = (int32_t *)calloc((size_t)nblocks * 64 + 1, 4);
      out->fdct = (int32_t *)calloc((size_t)nblocks * 64 + 1, 4);
    }
    i64 blk = 0, mcu = 0;
    int rst = 0;
    i64 fdct[64], quant[64], runs[65], values[65];
    for (int y_mb = 0; y_mb < mbs_high; y_mb++)
      for (int x_mb = 0; x_mb < mbs_wide; x_mb++, mcu++) {
        if (restart_interval > 0 && mcu > 0 && mcu % restart_interval == 0) { /* stated extension */
          orc_writer_flush_with_1s(w, 1);
          write_marker_code(w, M_RST0 + (rst & 7));
          rst++;
          for (int i = 0; i < p.ncomp; i++) scans[i].dc_pred = 0;
        }
        for (int i = 0; i < p.ncomp; i++) {
          enc_scan *s = &scans[i];
          for (int y_sub = 0; y_sub < s->vs; y_sub++)
            for (int x_sub = 0; x_sub < s->hs; x_sub++, blk++) {
              int x_pos = (x_mb * s->hs + x_sub) * 8, y_pos = (y_mb * s->vs + y_sub) * 8;
              /* level_shifted_input_block (:81-90), bounds-checked like Plane.![] */
              for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) {
                  if (x + x_pos >= s->pw || y + y_pos >= s->ph) orc_raise(ORC_ERR_PLANE_BOUNDS);
                  fdct[y * 8 + x] = (i64)s->plane[(size_t)(x + x_pos) + (size_t)(y + y_pos) * (size_t)s->pw] - 128;
                }
              orc_chen_forward_8x8(fdct); /* :92 */
              for (int k = 0; k < 64; k++) /* quant (:103-108) */
                quant[orc_zigzag_forward[k]] = quant_and_scale(fdct[k], s->quant_table[orc_zigzag_forward[k]]);
              int n = orc_rle(quant, &s->dc_pred, runs, values); /* rle (:127-141) */
              if (want_blocks)
                for (int k = 0; k < 64; k++) {
                  out->quant[blk * 64 + k] = (int32_t)quant[k];
                  out->fdct[blk * 64 + k] = (int32_t)fdct[k];
                }
              /* write_bits (:149-193) */
              {
                int size = orc_size(values[0]); /* write_dc */
                if (size >= s->n_dc) orc_raise(ORC_ERR_INVALID_ARG);
                orc_writer_put_bits(w, 1, s->dc_table[size].bits, s->dc_table[size].length);
                orc_writer_put_bits(w, 1, orc_magnitude(size, values[0]), size);
              }
              for (int e = 1; e < n; e++) {
                if (e == n - 1 && values[e] == 0) { /* [ {run; value = 0} ] -> end of block */
                  write_ac(w, s, 0, 0);
                } else {
                  i64 run = runs[e];
                  while (run >= 16) { /* ZRL */
                    write_ac(w, s, 15, 0);
                    run -= 16;
                  }
                  write_ac(w, s, (int)run, values[e]);
                }
              }
            }
        }
      }
    /* complete_and_write_eoi (:507-510) */
    orc_writer_flush_with_1s(w, 1);
    write_marker_code(w, M_EOI);
  }
  ORC_END_TRY;
  return st;
}

int orc_encode(const uint8_t *y, const uint8_t *u, const uint8_t *v, int width, int height, int chroma, int quality,
               int restart_interval, int want_blocks, orc_encoded *out) {
  memset(out, 0, sizeof(*out));
  orc_writer w;
  orc_writer_create(&w);
  enc_scan *scans = (enc_scan *)calloc(3, sizeof(enc_scan));
  const uint8_t *src[3] = {y, u, v};
  int st = encode_impl(src, width, height, chroma, quality, restart_interval, want_blocks, out, &w, scans);
  for (int i = 0; i < 3; i++) free(scans[i].plane);
  free(scans);
  out->status = st;
  if (st == ORC_OK) {
    out->bytes = w.buffer;
    out->len = w.bytes_written;
  } else {
    orc_writer_free(&w);
    orc_encoded_free(out);
    out->status = st;
  }
  return st;
}

void orc_encoded_free(orc_encoded *e) {
  free(e->bytes);
  free(e->quant);
  free(e->fdct);
  e->bytes = NULL;
  e->quant = e->fdct = NULL;
}

/* ------------------------------------------------------------------------------------------
 * tools/src: Planar_444, Yuv.crop, Ocompare
 * ---------------------------------------------------------------------------------------- */
static int avg2(int a, int b) { return (a + b + 1) >> 1; }                      /* planar_444.ml:4-8 */
static int avg4(int a, int b, int c, int d) { return (a + b + c + d + 2) >> 2; } /* :10-16 */

void orc_subsample_h2(const uint8_t *src, int w, int h, uint8_t *dst) { /* :18-23; w,h = src dims */
  int dw = w / 2;
  for (int row = 0; row < h; row++)
    for (int col = 0; col < dw; col++)
      dst[row * dw + col] = (uint8_t)avg2(src[row * w + col * 2], src[row * w + col * 2 + 1]);
}
void orc_supersample_h2(const uint8_t *src, int w, int h, uint8_t *dst) { /* :25-33; w,h = src dims */
  int dw = w * 2;
  for (int row = 0; row < h; row++) {
    for (int col = 0; col <= w - 2; col++) {
      dst[row * dw + col * 2] = src[row * w + col];
      dst[row * dw + col * 2 + 1] = (uint8_t)avg2(src[row * w + col], src[row * w + col + 1]);
    }
    dst[row * dw + w * 2 - 2] = src[row * w + w - 1];
    dst[row * dw + w * 2 - 1] = src[row * w + w - 1];
  }
}
void orc_subsample_hv2(const uint8_t *src, int w, int h, uint8_t *dst) { /* :68-80 */
  int dw = w / 2, dh = h / 2;
  for (int row = 0; row < dh; row++)
    for (int col = 0; col < dw; col++)
      dst[row * dw + col] = (uint8_t)avg4(src[(row * 2) * w + col * 2], src[(row * 2) * w + col * 2 + 1],
                                          src[(row * 2 + 1) * w + col * 2], src[(row * 2 + 1) * w + col * 2 + 1]);
}
void orc_supersample_hv2(const uint8_t *src, int w, int h, uint8_t *dst) { /* :82-103 */
  int dw = w * 2;
  for (int row = 0; row < h; row++) {
    int row1 = row, row2 = row + 1 < h - 1 ? row + 1 : h - 1;
    for (int col = 0; col <= w - 2; col++) {
      int a = src[row1 * w + col], b = src[row1 * w + col + 1], c = src[row2 * w + col], d = src[row2 * w + col + 1];
      dst[(row * 2) * dw + col * 2] = (uint8_t)a;
      dst[(row * 2) * dw + col * 2 + 1] = (uint8_t)avg2(a, b);
      dst[(row * 2 + 1) * dw + col * 2] = (uint8_t)avg2(a, c);
      dst[(row * 2 + 1) * dw + col * 2 + 1] = (uint8_t)avg4(a, b, c, d);
    }
    int a = src[row1 * w + w - 1], b = src[row2 * w + w - 1];
    dst[(row * 2) * dw + w * 2 - 2] = (uint8_t)a;
    dst[(row * 2) * dw + w * 2 - 1] = (uint8_t)a;
    dst[(row * 2 + 1) * dw + w * 2 - 2] = (uint8_t)avg2(a, b);
    dst[(row * 2 + 1) * dw + w * 2 - 1] = (uint8_t)avg2(a, b);
  }
}
void orc_crop_clamp(const uint8_t *src, int sw, int sh, int x_pos, int y_pos, uint8_t *dst, int dw, int dh) {
  for (int r = 0; r < dh; r++) /* yuv.ml:43-62 */
    for (int c = 0; c < dw; c++) {
      int col = c + x_pos, row = r + y_pos;
      col = col < 0 ? 0 : col >= sw ? sw - 1 : col;
      row = row < 0 ? 0 : row >= sh ? sh - 1 : row;
      dst[r * dw + c] = src[row * sw + col];
    }
}
i64 orc_square_error(const uint8_t *a, const uint8_t *b, i64 n) { /* ocompare.ml:41-52 */
  i64 acc = 0;
  for (i64 i = 0; i < n; i++) {
    i64 d = (i64)a[i] - (i64)b[i];
    acc += d * d;
  }
  return acc;
}
i64 orc_max_difference(const uint8_t *a, const uint8_t *b, i64 n) { /* ocompare.ml:8-17 */
  i64 m = 0;
  for (i64 i = 0; i < n; i++) {
    i64 d = (i64)a[i] - (i64)b[i];
    if (d < 0) d = -d;
    if (d > m) m = d;
  }
  return m;
}
double orc_psnr(const uint8_t *a, const uint8_t *b, int w, int h) { /* ocompare.ml:54-59 */
  double mse = (double)orc_square_error(a, b, (i64)w * h) / ((double)w * (double)h);
  return 10. * log10(255. * 255. / mse);
}

/* Stated formula (not in the reference): JFIF full-range BT.601, 16-bit fixed point, the constants
 * libjpeg's jdcolor.c uses: R = Y + 1.40200 Cr', G = Y - 0.34414 Cb' - 0.71414 Cr', B = Y + 1.77200 Cb'
 * with FIX(x) = (int)(x * 65536 + 0.5), ONE_HALF = 32768 and arithmetic right shifts. */
void orc_ycbcr_to_rgb24(const uint8_t *y, const uint8_t *cb, const uint8_t *cr, i64 n, uint8_t *rgb) {
  for (i64 i = 0; i < n; i++) {
    i64 Y = y[i], Cb = (i64)cb[i] - 128, Cr = (i64)cr[i] - 128;
    i64 r = Y + asr(91881 * Cr + 32768, 16);
    i64 g = Y + asr(-22554 * Cb - 46802 * Cr + 32768, 16);
    i64 b = Y + asr(116130 * Cb + 32768, 16);
    rgb[3 * i + 0] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
    rgb[3 * i + 1] = (uint8_t)(g < 0 ? 0 : g > 255 ? 255 : g);
    rgb[3 * i + 2] = (uint8_t)(b < 0 ? 0 : b > 255 ? 255 : b);
  }
}

/* ------------------------------------------------------------------------------------------
 * cpu_baseline timing helpers (bench.py)
 * ---------------------------------------------------------------------------------------- */
static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
double orc_time_decode(const uint8_t *const *jpeg, const i64 *len, int n, int flags, int reps) {
  double t0 = now_s();
  for (int r = 0; r < reps; r++)
    for (int i = 0; i < n; i++) {
      orc_decoded d;
      if (orc_decode(jpeg[i], len[i], flags, 0, &d) != ORC_OK) return -1.0;
      orc_decoded_free(&d);
    }
  return now_s() - t0;
}
double orc_time_encode(const uint8_t *yuv, int width, int height, int chroma, int quality, int restart_interval,
                       int reps) {
  int cw = chroma == 444 ? width : width / 2, ch = chroma == 420 ? height / 2 : height;
  const uint8_t *u = yuv + (size_t)width * height, *v = u + (size_t)cw * ch;
  double t0 = now_s();
  for (int r = 0; r < reps; r++) {
    orc_encoded e;
    if (orc_encode(yuv, u, v, width, height, chroma, quality, restart_interval, 0, &e) != ORC_OK) return -1.0;
    orc_encoded_free(&e);
  }
  return now_s() - t0;
}
