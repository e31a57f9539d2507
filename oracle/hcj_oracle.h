/*
 * hcj_oracle.h — CPU oracle for the hardcamls/video-coding JPEG software model.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * OCaml model (jpeg/model/src + common/src), kept deliberately literal so that
 * results can be compared bit-for-bit with the CUDA path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product library (libhcjpeg.so) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_goldens.py checks this code against
 * every golden the reference's own tests hold for the path (mini.jpg byte
 * identity, the 18 cram PSNR values, the Chen example block, the 623-byte
 * header, encoder code tables, quant scaling, size/magnitude, RLE cases,
 * Mouse480 header + entropy bytes, up-sampling vectors); see tests/golden/.
 * The OCaml model itself cannot be built here (no OCaml toolchain), so there
 * is no oracle/_ref.
 *
 * All arithmetic is int64_t ("int" in the 63-bit OCaml model); ">>" on signed
 * values is arithmetic (checked at start-up), "/" truncates toward zero.
 *
 * Citations "file.ml:LINE" are relative to the reference repository root.
 */
#ifndef HCJ_ORACLE_H
#define HCJ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes.  Same numeric values as include/hcjpeg.h (tests assert this). */
enum {
  ORC_OK = 0,
  ORC_ERR_UNSUPPORTED_MARKER = -1,  /* decoder.ml:67  "unsupported marker code" */
  ORC_ERR_NO_DC_CODE = -2,          /* decoder.ml:92  "Can't find dc code" */
  ORC_ERR_NO_AC_CODE = -3,          /* decoder.ml:101 "Can't find ac code" */
  ORC_ERR_COEF_INDEX = -4,          /* decoder.ml:136 "coefficient index out of range:" */
  ORC_ERR_NO_COMPONENT = -5,        /* decoder.ml:228 "unable to find component identifier" */
  ORC_ERR_NO_QUANT_TABLE = -6,      /* decoder.ml:234 "unable to find quantisation table" */
  ORC_ERR_NO_HUFFMAN_TABLE = -7,    /* decoder.ml:243 "unable to find huffman table" */
  ORC_ERR_NO_FRAME_OR_SCAN = -8,    /* decoder.ml:291 "From start of frame or start of scan marker" */
  ORC_ERR_BITS_OUT_OF_BOUNDS = -9,  /* bitstream_reader.ml:32 "Bitstream_reader out of bounds" */
  ORC_ERR_PLANE_BOUNDS = -10,       /* plane.ml:47-59 "[Plane.get/set] out of bounds" */
  ORC_ERR_FRAME_INFER = -11,        /* frame.ml:44,55 chroma planes mismatch / cannot infer */
  ORC_ERR_NEED_3_COMPONENTS = -12,  /* decoder.ml:415-420 components.(1)/.(2) index out of bounds */
  ORC_ERR_ENCODER_PARAMS = -13,     /* encoder.ml:274-282, model.ml:66-70 */
  /* Stated extensions / places where the reference does not terminate: */
  ORC_ERR_NO_TERMINATOR = -20,      /* decoder.ml:261-281 loops forever without a marker after the scan */
  ORC_ERR_RESTART_COUNT = -21,      /* restart extension: RSTn count != ceil(MCUs/Ri)-1 */
  ORC_ERR_UNSUPPORTED_GEOMETRY = -22, /* >4 scan components, sampling factor 0 or >4, >10 blocks/MCU, DC category >15 */
  ORC_ERR_DC_RANGE = -23,           /* absolute DC does not fit int16 in the coefficient tap */
  ORC_ERR_TRUNCATED = -24,          /* decoder.ml:24-29 find_marker never returns on a truncated header */
  ORC_ERR_BAD_HUFFMAN_TABLE = -25,  /* tables.ml:497-499 array index out of bounds (over-subscribed DHT) */
  ORC_ERR_BUFFER_TOO_SMALL = -30,
  ORC_ERR_INVALID_ARG = -31
};

/* ---- Bitstream_reader.From_string (common/src/bitstream_reader.ml:6-57) ---- */
typedef struct {
  const uint8_t *buf;
  int64_t len;            /* bytes */
  int64_t length_in_bits; /* :16 */
  int64_t bit_pos;
} orc_bits;

void orc_bits_create(orc_bits *b, const uint8_t *buf, int64_t len);
int orc_bits_show(orc_bits *b, int n, int64_t *v); /* status; raises iff n >= length_in_bits (:32) */
int orc_bits_get(orc_bits *b, int n, int64_t *v);
void orc_bits_advance(orc_bits *b, int64_t n);
void orc_bits_align_to_byte(orc_bits *b);

/* ---- Bitstream_writer (common/src/bitstream_writer.ml:3-49) ---- */
typedef struct {
  uint64_t word_buffer;
  int word_bits;
  uint8_t *buffer;
  int64_t bytes_written;
  int64_t capacity;
} orc_writer;

void orc_writer_create(orc_writer *w);
void orc_writer_free(orc_writer *w);
void orc_writer_put_bits(orc_writer *w, int stuffing, int64_t value, int bits);
void orc_writer_flush_with_1s(orc_writer *w, int stuffing);

/* ---- Markers (jpeg/model/src/markers.ml) / Decoder.Header (decoder.ml:5-71) ---- */
#define ORC_MAX_COMPONENTS 255
#define ORC_MAX_TABLES 64 /* DQT/DHT segments remembered (list order = newest first) */

typedef struct {
  int identifier, horizontal_sampling_factor, vertical_sampling_factor, quantization_table_identifier;
} orc_component;

typedef struct {
  int present;
  int length, sample_precision, width, height, number_of_components;
  orc_component components[ORC_MAX_COMPONENTS];
} orc_sof;

typedef struct {
  int selector, dc_coef_selector, ac_coef_selector;
} orc_scan_component;

typedef struct {
  int present;
  int length, number_of_image_components;
  orc_scan_component scan_components[ORC_MAX_COMPONENTS];
  int start_of_predictor_selection, end_of_predictor_selection;
  int successive_approximation_bit_high, successive_approximation_bit_low;
} orc_sos;

typedef struct {
  int length, element_precision, table_identifier;
  int64_t elements[64];
} orc_dqt;

typedef struct {
  int length, table_class, destination_identifier;
  int lengths[16];
  int nvalues;
  int values[16 * 255];
} orc_dht;

typedef struct {
  orc_sof frame;
  orc_sos scan;
  int restart_interval_present;
  int restart_interval_length, restart_interval;
  int n_quant_tables;   /* quant_tables[0] is the most recently parsed (list head) */
  orc_dqt quant_tables[ORC_MAX_TABLES];
  int n_huffman_tables; /* likewise */
  orc_dht huffman_tables[ORC_MAX_TABLES];
  int64_t scan_bit_pos; /* reader position after SOS (first entropy-coded byte * 8) */
} orc_header;

int orc_header_decode(const uint8_t *jpeg, int64_t len, orc_header *h);
int orc_header_decode_ex(const uint8_t *jpeg, int64_t len, int flags, orc_header *h);

/* For_testing.extract_entropy_coded_bits (decoder.ml:261-281).  out must hold len bytes. */
int orc_extract_entropy_coded_bits(const uint8_t *jpeg, int64_t len, int64_t start_byte,
                                   uint8_t *out, int64_t *out_len);

/* ---- Tables (jpeg/model/src/tables.ml) ---- */
typedef struct {
  int length, bits, data; /* data: dc category, or (run<<4)|size for ac */
} orc_code;

/* Specification.create_code_table (tables.ml:27-45); codes[] needs room for sum(lengths). */
int orc_create_code_table(const int lengths[16], const int *values, orc_code *codes);
/* Default specs (tables.ml:54-476): which = 0 dc_luma, 1 dc_chroma, 2 ac_luma, 3 ac_chroma */
void orc_default_spec(int which, const int **lengths, const int **values, int *nvalues);
/* Encoder.dc_table / ac_table (tables.ml:504-545): ac is [16][11] with zero-length dummies. */
int orc_encoder_dc_table(int which, orc_code *out /*[16]*/, int *n);
int orc_encoder_ac_table(int which, orc_code *out /*[16*16]*/, int row_len[16]);

/* ---- Quant_tables / Zigzag ---- */
void orc_quant_scale(int chroma, int quality, int64_t out[64]); /* quant_tables.ml:139-147 */
extern const int orc_zigzag_inverse[64];                        /* zigzag.ml:3-69   */
extern const int orc_zigzag_forward[64];                        /* zigzag.ml:71-137 */

/* ---- Dct.Chen (jpeg/model/src/dct.ml:3-197) ---- */
void orc_chen_inverse_8x8(int64_t block[64]);
void orc_chen_forward_8x8(int64_t block[64]);

/* ---- codewords (encoder.ml:143-147, decoder.ml:73-79) ---- */
int orc_size(int64_t value);
int64_t orc_magnitude(int size, int64_t value);
int64_t orc_mag(int cat, int64_t code);
/* Encoder.rle (encoder.ml:127-141): returns number of (run,value) pairs written. */
int orc_rle(const int64_t quant[64], int64_t *dc_pred, int64_t runs[65], int64_t values[65]);

/* ---- Decoder (jpeg/model/src/decoder.ml) ---- */
#define ORC_FLAG_RESTART_EXT 1 /* stated extension: honour DRI/RSTn (T.81 semantics) */
#define ORC_FLAG_T81_TABLES 2  /* stated extension: several tables per DQT/DHT segment, FF fill bytes before markers */

typedef struct {
  int status;
  int ncomp;
  int width, height;             /* frame */
  int mcus_wide, mcus_high, blocks_per_mcu;
  int64_t nblocks;
  /* per scan component (decoder.ml:167-187) */
  int hs[4], vs[4];
  int decoded_width[4], decoded_height[4], actual_width[4], actual_height[4];
  uint8_t *plane[4];             /* padded planes (get_decoded_planes, :399-401) */
  uint8_t *cropped[4];           /* crop (:403-413) */
  /* per block, traversal order of decode_seq (:374-395) */
  int32_t *coefs;                /* [nblocks][64] zig-zag, coefs[0] = DC differential (Component.coefs) */
  int32_t *dc_abs;               /* [nblocks] dc_pred after the block */
  int32_t *dequant;              /* [nblocks][64] natural order (Component.dequant) */
  uint8_t *recon;                /* [nblocks][64] (Component.recon) */
  int8_t *block_comp;            /* [nblocks] scan-component index */
  int64_t entropy_len;           /* destuffed bytes (all intervals) */
  int yuv_status;                /* status of get_yuv_frame (Frame.of_planes) */
  int chroma;                    /* 420 / 422 / 444 when yuv_status == 0 */
} orc_decoded;

/* Header.decode + init + decode (+ get_yuv_frame).  want_blocks != 0 also fills the per-block taps. */
int orc_decode(const uint8_t *jpeg, int64_t len, int flags, int want_blocks, orc_decoded *out);
void orc_decoded_free(orc_decoded *d);

/* ---- Encoder (jpeg/model/src/encoder.ml) ---- */
typedef struct {
  int status;
  uint8_t *bytes; /* Writer.get_buffer */
  int64_t len;
  int64_t nblocks;
  int32_t *quant; /* [nblocks][64] zig-zag (Block.quant), optional */
  int32_t *fdct;  /* [nblocks][64] natural (Block.fdct), optional */
} orc_encoded;

/* chroma: 420 / 422 / 444 / 400 (encode_monochrome).  restart_interval > 0 is the stated
 * extension (DRI + RSTn); 0 reproduces the reference encoder exactly. */
int orc_encode(const uint8_t *y, const uint8_t *u, const uint8_t *v, int width, int height, int chroma,
               int quality, int restart_interval, int want_blocks, orc_encoded *out);
void orc_encoded_free(orc_encoded *e);
int orc_write_headers(int width, int height, int chroma, int quality, int restart_interval,
                      uint8_t *out, int64_t cap, int64_t *len);

/* ---- tools: Planar_444 (tools/src/planar_444.ml), Ocompare (tools/src/ocompare.ml), Yuv.crop ---- */
void orc_supersample_h2(const uint8_t *src, int w, int h, uint8_t *dst);  /* :25-33  */
void orc_supersample_hv2(const uint8_t *src, int w, int h, uint8_t *dst); /* :82-103 */
void orc_subsample_h2(const uint8_t *src, int w, int h, uint8_t *dst);    /* :18-23  */
void orc_subsample_hv2(const uint8_t *src, int w, int h, uint8_t *dst);   /* :68-80  */
void orc_crop_clamp(const uint8_t *src, int sw, int sh, int x_pos, int y_pos, uint8_t *dst, int dw, int dh); /* yuv.ml:43-62 */
int64_t orc_square_error(const uint8_t *a, const uint8_t *b, int64_t n);  /* ocompare.ml:41-52 */
int64_t orc_max_difference(const uint8_t *a, const uint8_t *b, int64_t n);/* ocompare.ml:8-17 */
double orc_psnr(const uint8_t *a, const uint8_t *b, int w, int h);        /* ocompare.ml:54-59 */

/* YCbCr -> RGB24.  NOT in the reference ("parity unpinned"): stated formula, JFIF full range,
 * 16-bit fixed point, see DESIGN.md.  Inputs are 4:4:4 planes. */
void orc_ycbcr_to_rgb24(const uint8_t *y, const uint8_t *cb, const uint8_t *cr, int64_t n, uint8_t *rgb);

/* Timing helper for bench.py's cpu_baseline: decode `n` images `reps` times, return seconds. */
double orc_time_decode(const uint8_t *const *jpeg, const int64_t *len, int n, int flags, int reps);
double orc_time_encode(const uint8_t *yuv, int width, int height, int chroma, int quality,
                       int restart_interval, int reps);

#ifdef __cplusplus
}
#endif
#endif
