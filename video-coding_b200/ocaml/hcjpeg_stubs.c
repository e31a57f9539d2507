/* hcjpeg_stubs.c — thin OCaml `external` stubs over the C ABI of libhcjpeg (include/hcjpeg.h).
 *
 * Mechanical by design (it cannot be compiled in this repository's image: no caml/ headers; reviewed against
 * caml/mlvalues.h conventions).  Inputs are OCaml strings / Bigarrays; strings are copied before the runtime lock is
 * released (they may move), Bigarray data does not move.  Outputs are freshly allocated Bigarrays (C layout, char =
 * Base_bigstring.t) or OCaml strings.  A non-zero hcj_status becomes Failure with hcj_strerror's text, which is the
 * text of the model's own raise_s message for that condition.
 *
 * Contexts: one hcj_ctx per visible GPU, created on first use.  hcj_ctx is not thread-safe and the stubs release
 * the runtime lock around device work, so every use of the contexts is serialised by one mutex (two OCaml threads
 * calling in queue up; a single call already spreads a batch over all GPUs through hcj_decode_batch_multi). */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>
#include <caml/threads.h>

#include "hcjpeg.h"

#define HCJ_ML_MAX_GPUS 16
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static hcj_ctx *g_ctx[HCJ_ML_MAX_GPUS];
static int g_nctx = 0;

/* Called with the runtime lock held and g_lock NOT held.  Returns 0 or the failing status. */
static int ensure_contexts(void) {
  int st = HCJ_OK;
  pthread_mutex_lock(&g_lock);
  if (g_nctx == 0) {
    int n = hcj_device_count();
    const char *lim = getenv("HCJPEG_GPUS"); /* optional cap on the devices used */
    if (lim && atoi(lim) > 0 && atoi(lim) < n) n = atoi(lim);
    if (n > HCJ_ML_MAX_GPUS) n = HCJ_ML_MAX_GPUS;
    if (n < 1) n = 1; /* hcj_ctx_create reports the CUDA error: there is no CPU fallback */
    for (int k = 0; k < n && st == HCJ_OK; k++) {
      st = hcj_ctx_create(k, NULL, &g_ctx[k]);
      if (st == HCJ_OK) g_nctx = k + 1;
    }
    if (st != HCJ_OK) {
      for (int k = 0; k < g_nctx; k++) hcj_ctx_destroy(g_ctx[k]);
      g_nctx = 0;
    }
  }
  pthread_mutex_unlock(&g_lock);
  return st;
}

static void check(int st) {
  if (st != HCJ_OK) caml_failwith(hcj_strerror(st));
}

static size_t out_bytes_of(const hcj_frame_info *f, int mode) {
  return mode == HCJ_OUT_YUV ? f->yuv_bytes : mode == HCJ_OUT_PLANES ? f->planes_bytes : f->rgb_bytes;
}

/* external device_count : unit -> int */
CAMLprim value hcj_ml_device_count(value v_unit) {
  (void)v_unit;
  return Val_int(hcj_device_count());
}

/* external frame_info : string -> int array
   [| width; height; ncomp; chroma; mcus_wide; mcus_high; blocks_per_mcu;
      decoded_w.(0..3); decoded_h.(0..3); actual_w.(0..3); actual_h.(0..3); nblocks |] */
CAMLprim value hcj_ml_frame_info(value v_bits) {
  CAMLparam1(v_bits);
  CAMLlocal1(v_res);
  hcj_frame_info f;
  check(hcj_frame_info_get((const uint8_t *)String_val(v_bits), caml_string_length(v_bits), &f));
  v_res = caml_alloc_tuple(7 + 16 + 1);
  int k = 0;
  Store_field(v_res, k++, Val_int(f.width));
  Store_field(v_res, k++, Val_int(f.height));
  Store_field(v_res, k++, Val_int(f.ncomp));
  Store_field(v_res, k++, Val_int(f.chroma));
  Store_field(v_res, k++, Val_int(f.mcus_wide));
  Store_field(v_res, k++, Val_int(f.mcus_high));
  Store_field(v_res, k++, Val_int(f.blocks_per_mcu));
  for (int i = 0; i < 4; i++) Store_field(v_res, k++, Val_int(f.decoded_width[i]));
  for (int i = 0; i < 4; i++) Store_field(v_res, k++, Val_int(f.decoded_height[i]));
  for (int i = 0; i < 4; i++) Store_field(v_res, k++, Val_int(f.actual_width[i]));
  for (int i = 0; i < 4; i++) Store_field(v_res, k++, Val_int(f.actual_height[i]));
  Store_field(v_res, k++, Val_long((long)f.nblocks));
  CAMLreturn(v_res);
}

/* external header_decode : string -> int array        (Decoder.Header.decode, decoder.ml:37-70; host only)
   Flat encoding, decoded by Hcjpeg_gpu.Header.of_raw:
     has_frame; sof_length; sample_precision; width; height; number_of_components; 4 x (identifier; h; v; tq);
     has_scan; sos_length; number_of_image_components; 4 x (selector; dc; ac); ss; se; ah; al;
     has_restart_interval; dri_length; restart_interval; scan_byte_pos; n_quant_tables; n_huffman_tables;
     per quant table (list order):   length; element_precision; table_identifier; 64 elements
     per huffman table (list order): length; table_class; destination_identifier; 16 lengths; nvalues; values */
CAMLprim value hcj_ml_header_decode(value v_bits) {
  CAMLparam1(v_bits);
  CAMLlocal1(v_res);
  hcj_header *h = (hcj_header *)caml_stat_alloc(sizeof(hcj_header));
  int st = hcj_header_decode((const uint8_t *)String_val(v_bits), caml_string_length(v_bits), h);
  if (st != HCJ_OK) {
    caml_stat_free(h);
    check(st);
  }
  size_t n = 6 + 16 + 3 + 12 + 4 + 3 + 1 + 2 + (size_t)h->n_quant_tables * 67;
  for (int i = 0; i < h->n_huffman_tables; i++) n += 3 + 16 + 1 + (size_t)h->huffman_tables[i].nvalues;
  v_res = caml_alloc(n, 0); /* may be larger than Max_young_wosize: caml_alloc + caml_initialize-free stores of immediates */
  size_t k = 0;
#define PUT(x) Store_field(v_res, k++, Val_long((long)(x)))
  PUT(h->has_frame); PUT(h->sof_length); PUT(h->sample_precision); PUT(h->width); PUT(h->height); PUT(h->number_of_components);
  for (int i = 0; i < 4; i++) {
    PUT(h->components[i].identifier); PUT(h->components[i].horizontal_sampling_factor);
    PUT(h->components[i].vertical_sampling_factor); PUT(h->components[i].quantization_table_identifier);
  }
  PUT(h->has_scan); PUT(h->sos_length); PUT(h->number_of_image_components);
  for (int i = 0; i < 4; i++) {
    PUT(h->scan_components[i].selector); PUT(h->scan_components[i].dc_coef_selector); PUT(h->scan_components[i].ac_coef_selector);
  }
  PUT(h->start_of_predictor_selection); PUT(h->end_of_predictor_selection);
  PUT(h->successive_approximation_bit_high); PUT(h->successive_approximation_bit_low);
  PUT(h->has_restart_interval); PUT(h->dri_length); PUT(h->restart_interval);
  PUT(h->scan_byte_pos); PUT(h->n_quant_tables); PUT(h->n_huffman_tables);
  for (int i = 0; i < h->n_quant_tables; i++) {
    const hcj_dqt *q = &h->quant_tables[i];
    PUT(q->length); PUT(q->element_precision); PUT(q->table_identifier);
    for (int e = 0; e < 64; e++) PUT(q->elements[e]);
  }
  for (int i = 0; i < h->n_huffman_tables; i++) {
    const hcj_dht *t = &h->huffman_tables[i];
    PUT(t->length); PUT(t->table_class); PUT(t->destination_identifier);
    for (int e = 0; e < 16; e++) PUT(t->lengths[e]);
    PUT(t->nvalues);
    for (int e = 0; e < t->nvalues; e++) PUT(t->values[e]);
  }
#undef PUT
  caml_stat_free(h);
  CAMLreturn(v_res);
}

/* external decode_batch : string array -> int -> Bigstring.t array
   Decoder.decode_a_frame (decoder.ml:422-427) for every file, spread over all GPUs (hcj_decode_batch_multi).
   mode 0: cropped planar Y,U,V; 1: padded planes (get_decoded_planes); 2: RGB24; 3: planar 4:4:4.
   Raises Failure for the first image whose status is not HCJ_OK (the model raises at the first bad image too). */
CAMLprim value hcj_ml_decode_batch(value v_files, value v_mode) {
  CAMLparam2(v_files, v_mode);
  CAMLlocal2(v_res, v_ba);
  const int mode = Int_val(v_mode);
  const int n = (int)Wosize_val(v_files);
  check(ensure_contexts());
  const uint8_t **in = (const uint8_t **)caml_stat_alloc(sizeof(*in) * (n ? n : 1));
  uint8_t **out = (uint8_t **)caml_stat_alloc(sizeof(*out) * (n ? n : 1));
  size_t *len = (size_t *)caml_stat_alloc(sizeof(*len) * (n ? n : 1));
  size_t *cap = (size_t *)caml_stat_alloc(sizeof(*cap) * (n ? n : 1));
  int *status = (int *)caml_stat_alloc(sizeof(*status) * (n ? n : 1));
  v_res = caml_alloc(n, 0);
  int bad = HCJ_OK;
  for (int i = 0; i < n; i++) {
    value s = Field(v_files, i);
    len[i] = caml_string_length(s);
    hcj_frame_info f;
    int st = hcj_frame_info_get((const uint8_t *)String_val(s), len[i], &f);
    cap[i] = st == HCJ_OK ? out_bytes_of(&f, mode) : 0;
    if (st != HCJ_OK && bad == HCJ_OK) bad = st;
    intnat dims[1] = {(intnat)cap[i]};
    v_ba = caml_ba_alloc(CAML_BA_CHAR | CAML_BA_C_LAYOUT, 1, NULL, dims);
    Store_field(v_res, i, v_ba);
    out[i] = (uint8_t *)Caml_ba_data_val(v_ba);
  }
  /* OCaml strings may move once the lock is released: work from private copies (taken after the last allocation) */
  for (int i = 0; i < n; i++) {
    uint8_t *copy = (uint8_t *)caml_stat_alloc(len[i] ? len[i] : 1);
    memcpy(copy, String_val(Field(v_files, i)), len[i]);
    in[i] = copy;
  }
  int st = HCJ_OK;
  if (bad == HCJ_OK && n > 0) {
    caml_release_runtime_system();
    pthread_mutex_lock(&g_lock);
    st = hcj_decode_batch_multi(g_ctx, n < g_nctx ? n : g_nctx, in, len, n, mode, HCJ_FLAG_DEFAULT, out, cap, status);
    pthread_mutex_unlock(&g_lock);
    caml_acquire_runtime_system();
    for (int i = 0; i < n && st == HCJ_OK && bad == HCJ_OK; i++) bad = status[i];
  }
  for (int i = 0; i < n; i++) caml_stat_free((void *)in[i]);
  caml_stat_free(in);
  caml_stat_free(out);
  caml_stat_free(len);
  caml_stat_free(cap);
  caml_stat_free(status);
  check(st);
  check(bad);
  CAMLreturn(v_res);
}

/* external encode_batch : Bigstring.t array -> int -> int -> int -> int -> string array
   Encoder.encode_420 / 422 / 444 / monochrome (encoder.ml:522-552; chroma = 420 / 422 / 444 / 400) for frames of one
   geometry, spread over all GPUs.  Each frame is planar Y,U,V exactly as Frame.output writes it.  Output buffers are
   sized by hcj_encode_bound (the bound that never fails; a guess such as 3 bytes per pixel does). */
CAMLprim value hcj_ml_encode_batch(value v_frames, value v_w, value v_h, value v_chroma, value v_quality) {
  CAMLparam5(v_frames, v_w, v_h, v_chroma, v_quality);
  CAMLlocal2(v_res, v_str);
  const int w = Int_val(v_w), h = Int_val(v_h), chroma = Int_val(v_chroma), q = Int_val(v_quality);
  const int n = (int)Wosize_val(v_frames);
  check(ensure_contexts());
  const size_t bound = hcj_encode_bound(w, h, chroma);
  if (bound == 0) check(HCJ_ERR_ENCODER_PARAMS);
  const uint8_t **in = (const uint8_t **)caml_stat_alloc(sizeof(*in) * (n ? n : 1));
  uint8_t **out = (uint8_t **)caml_stat_alloc(sizeof(*out) * (n ? n : 1));
  size_t *cap = (size_t *)caml_stat_alloc(sizeof(*cap) * (n ? n : 1));
  size_t *len = (size_t *)caml_stat_alloc(sizeof(*len) * (n ? n : 1));
  int *status = (int *)caml_stat_alloc(sizeof(*status) * (n ? n : 1));
  for (int i = 0; i < n; i++) {
    in[i] = (const uint8_t *)Caml_ba_data_val(Field(v_frames, i)); /* bigarray data does not move */
    out[i] = (uint8_t *)caml_stat_alloc(bound);
    cap[i] = bound;
    len[i] = 0;
    status[i] = HCJ_OK;
  }
  int st = HCJ_OK;
  if (n > 0) {
    caml_release_runtime_system();
    pthread_mutex_lock(&g_lock);
    st = hcj_encode_batch_multi(g_ctx, n < g_nctx ? n : g_nctx, in, n, w, h, chroma, q, 0, out, cap, len, status);
    pthread_mutex_unlock(&g_lock);
    caml_acquire_runtime_system();
  }
  int bad = st;
  for (int i = 0; i < n && bad == HCJ_OK; i++) bad = status[i];
  v_res = caml_alloc(n, 0);
  for (int i = 0; i < n && bad == HCJ_OK; i++) {
    v_str = caml_alloc_initialized_string(len[i], (const char *)out[i]);
    Store_field(v_res, i, v_str);
  }
  for (int i = 0; i < n; i++) caml_stat_free(out[i]);
  caml_stat_free(in);
  caml_stat_free(out);
  caml_stat_free(cap);
  caml_stat_free(len);
  caml_stat_free(status);
  check(bad);
  CAMLreturn(v_res);
}

/* external decode_log : string -> Bigstring.t
   `model decode log` (jpeg/bin/model.ml:46-68): hcj_block_log records (include/hcjpeg.h) of every block in decode_seq
   order, packed back to back; Hcjpeg_gpu.Decoder.decode_log unpacks them. */
CAMLprim value hcj_ml_decode_log(value v_bits) {
  CAMLparam1(v_bits);
  CAMLlocal1(v_out);
  check(ensure_contexts());
  const size_t len = caml_string_length(v_bits);
  hcj_frame_info f;
  check(hcj_frame_info_get((const uint8_t *)String_val(v_bits), len, &f));
  intnat dims[1] = {(intnat)((size_t)f.nblocks * sizeof(hcj_block_log))};
  v_out = caml_ba_alloc(CAML_BA_CHAR | CAML_BA_C_LAYOUT, 1, NULL, dims);
  uint8_t *copy = (uint8_t *)caml_stat_alloc(len ? len : 1);
  memcpy(copy, String_val(v_bits), len);
  hcj_block_log *out = (hcj_block_log *)Caml_ba_data_val(v_out);
  const uint8_t *in[1] = {copy};
  int status[1] = {HCJ_OK};
  caml_release_runtime_system();
  pthread_mutex_lock(&g_lock);
  hcj_batch *b = NULL;
  int st = hcj_batch_create(g_ctx[0], in, &len, 1, HCJ_OUT_PLANES, HCJ_FLAG_DEFAULT, status, &b);
  if (st == HCJ_OK && status[0] == HCJ_OK) st = hcj_batch_decode(g_ctx[0], b);
  if (st == HCJ_OK && status[0] == HCJ_OK) st = hcj_batch_fetch_block_log(g_ctx[0], b, 0, 0, (size_t)f.nblocks, out);
  if (b) hcj_batch_destroy(g_ctx[0], b);
  pthread_mutex_unlock(&g_lock);
  caml_acquire_runtime_system();
  caml_stat_free(copy);
  check(st);
  check(status[0]);
  CAMLreturn(v_out);
}

/* external encode_log : Bigstring.t -> int -> int -> int -> int -> Bigstring.t
   `model encode log -verbose` (jpeg/bin/model.ml:108-142): hcj_encoder_block records of every block in encode_seq order. */
CAMLprim value hcj_ml_encode_log(value v_yuv, value v_w, value v_h, value v_chroma, value v_quality) {
  CAMLparam5(v_yuv, v_w, v_h, v_chroma, v_quality);
  CAMLlocal1(v_out);
  const int w = Int_val(v_w), h = Int_val(v_h), chroma = Int_val(v_chroma), q = Int_val(v_quality);
  check(ensure_contexts());
  uint8_t hdr[1024];
  size_t hlen = 0;
  check(hcj_write_headers(w, h, chroma, q, 0, hdr, sizeof(hdr) - 2, &hlen));
  hdr[hlen] = 0xff; /* a header followed by EOI is enough for the geometry */
  hdr[hlen + 1] = 0xd9;
  hcj_frame_info f;
  check(hcj_frame_info_get(hdr, hlen + 2, &f));
  intnat dims[1] = {(intnat)((size_t)f.nblocks * sizeof(hcj_encoder_block))};
  v_out = caml_ba_alloc(CAML_BA_CHAR | CAML_BA_C_LAYOUT, 1, NULL, dims);
  const uint8_t *yuv = (const uint8_t *)Caml_ba_data_val(v_yuv);
  hcj_encoder_block *out = (hcj_encoder_block *)Caml_ba_data_val(v_out);
  caml_release_runtime_system();
  pthread_mutex_lock(&g_lock);
  int st = hcj_encode_block_log(g_ctx[0], yuv, w, h, chroma, q, 0, 0, (size_t)f.nblocks, out);
  pthread_mutex_unlock(&g_lock);
  caml_acquire_runtime_system();
  check(st);
  CAMLreturn(v_out);
}
