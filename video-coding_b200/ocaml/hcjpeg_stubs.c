/* hcjpeg_stubs.c — thin OCaml `external` stubs over the C ABI of libhcjpeg (include/hcjpeg.h).
 *
 * Mechanical by design (it cannot be compiled in this repository's image: no caml/ headers).
 * Conventions: inputs are OCaml strings (copied by the library into device staging before the runtime
 * lock is released), outputs are freshly allocated Bigarrays (C layout, char) that Base_bigstring /
 * Plane.t wrap without a copy.  A non-zero hcj_status is turned into Failure with hcj_strerror's text,
 * which is the text of the model's own raise_s message for that condition.
 */
#include <string.h>

#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>
#include <caml/threads.h>

#include "hcjpeg.h"

static hcj_ctx *the_ctx(void) {
  static hcj_ctx *ctx = NULL;
  if (!ctx) {
    int st = hcj_ctx_create(0, NULL, &ctx);
    if (st != HCJ_OK) caml_failwith(hcj_strerror(st));
  }
  return ctx;
}

static void check(int st) {
  if (st != HCJ_OK) caml_failwith(hcj_strerror(st));
}

/* external frame_info : string -> int array
   [| width; height; ncomp; chroma; mcus_wide; mcus_high; blocks_per_mcu;
      decoded_w.(0..3); decoded_h.(0..3); actual_w.(0..3); actual_h.(0..3) |] */
CAMLprim value hcj_ml_frame_info(value v_bits) {
  CAMLparam1(v_bits);
  CAMLlocal1(v_res);
  hcj_frame_info f;
  check(hcj_frame_info_get((const uint8_t *)String_val(v_bits), caml_string_length(v_bits), &f));
  v_res = caml_alloc_tuple(7 + 16);
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
  CAMLreturn(v_res);
}

/* external decode : string -> int -> (char, int8_unsigned_elt, c_layout) Bigarray.Array1.t
   mode 0: cropped planar Y,U,V (Decoder.decode_a_frame); 1: padded planes (get_decoded_planes); 2: RGB24 */
CAMLprim value hcj_ml_decode(value v_bits, value v_mode) {
  CAMLparam2(v_bits, v_mode);
  CAMLlocal1(v_out);
  const int mode = Int_val(v_mode);
  const size_t len = caml_string_length(v_bits);
  hcj_frame_info f;
  check(hcj_frame_info_get((const uint8_t *)String_val(v_bits), len, &f));
  size_t bytes = mode == HCJ_OUT_YUV ? f.yuv_bytes : mode == HCJ_OUT_PLANES ? f.planes_bytes : f.rgb_bytes;
  intnat dims[1] = {(intnat)bytes};
  v_out = caml_ba_alloc(CAML_BA_CHAR | CAML_BA_C_LAYOUT, 1, NULL, dims);
  /* The OCaml string may move once the lock is released: work from a private copy. */
  uint8_t *copy = (uint8_t *)caml_stat_alloc(len ? len : 1);
  memcpy(copy, String_val(v_bits), len);
  const uint8_t *in[1] = {copy};
  uint8_t *out[1] = {(uint8_t *)Caml_ba_data_val(v_out)};
  size_t cap[1] = {bytes};
  int status[1] = {0};
  hcj_ctx *ctx = the_ctx();
  caml_release_runtime_system();
  int st = hcj_decode_batch(ctx, in, &len, 1, mode, HCJ_FLAG_DEFAULT, out, cap, status);
  caml_acquire_runtime_system();
  caml_stat_free(copy);
  check(st);
  check(status[0]);
  CAMLreturn(v_out);
}

/* external encode : Bigstring.t -> width:int -> height:int -> chroma:int -> quality:int -> string
   The frame is planar Y,U,V exactly as Frame.output writes it. */
CAMLprim value hcj_ml_encode(value v_yuv, value v_w, value v_h, value v_chroma, value v_quality) {
  CAMLparam5(v_yuv, v_w, v_h, v_chroma, v_quality);
  CAMLlocal1(v_out);
  const int w = Int_val(v_w), h = Int_val(v_h), chroma = Int_val(v_chroma), q = Int_val(v_quality);
  size_t cap = (size_t)w * h * 3 + 65536, len = 0;
  uint8_t *buf = (uint8_t *)caml_stat_alloc(cap);
  const uint8_t *in[1] = {(const uint8_t *)Caml_ba_data_val(v_yuv)}; /* bigarray data does not move */
  uint8_t *out[1] = {buf};
  int status[1] = {0};
  hcj_ctx *ctx = the_ctx();
  caml_release_runtime_system();
  int st = hcj_encode_batch(ctx, in, 1, w, h, chroma, q, 0, out, &cap, &len, status);
  caml_acquire_runtime_system();
  if (st != HCJ_OK || status[0] != HCJ_OK) {
    caml_stat_free(buf);
    check(st);
    check(status[0]);
  }
  v_out = caml_alloc_initialized_string(len, (const char *)buf);
  caml_stat_free(buf);
  CAMLreturn(v_out);
}
