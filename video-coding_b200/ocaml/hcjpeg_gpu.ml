open Base
open Hardcaml_video_common
open Hardcaml_jpeg_model
module Bigstring = Base_bigstring
module Model_decoder = Decoder
module Model_encoder = Encoder

external device_count : unit -> int = "hcj_ml_device_count"
external frame_info : string -> int array = "hcj_ml_frame_info"
external header_decode_raw : string -> int array = "hcj_ml_header_decode"
external decode_batch : string array -> int -> Bigstring.t array = "hcj_ml_decode_batch"

external encode_batch
  :  Bigstring.t array
  -> int
  -> int
  -> int
  -> int
  -> string array
  = "hcj_ml_encode_batch"

external decode_log_raw : string -> Bigstring.t = "hcj_ml_decode_log"

external encode_log_raw
  :  Bigstring.t
  -> int
  -> int
  -> int
  -> int
  -> Bigstring.t
  = "hcj_ml_encode_log"

(* ---- Plane.t <-> bytes, through the interface plane.mli exports (Plane.t is abstract: plane.mli:6) ---- *)

(* [width * height] bytes of [buf] at [pos] as a new plane: Plane.create (plane.mli:9) + the 1-d setter (plane.mli:26).
   With the optional patch in patches/plane_of_bigstring.diff this is a single Bigstring.blit. *)
let plane_of buf ~pos ~width ~height =
  let p = Plane.create ~width ~height in
  for i = 0 to (width * height) - 1 do
    Plane.(p.!(i) <- Bigstring.get buf (pos + i))
  done;
  p
;;

(* The bytes of [p] into [buf] at [pos] (1-d getter, plane.mli:25); returns the position behind them. *)
let plane_into p buf ~pos =
  let n = Plane.width p * Plane.height p in
  for i = 0 to n - 1 do
    Bigstring.set buf (pos + i) Plane.(p.!(i))
  done;
  pos + n
;;

(* The file a reader was created from, from its current (byte-aligned) position: Bits.get_buffer / Bits.bit_pos
   (bitstream_reader_intf.ml:13,27). *)
let file_of bits =
  let s = Model_decoder.Bits.get_buffer bits in
  let pos = (Model_decoder.Bits.bit_pos bits + 7) / 8 in
  if pos = 0 then s else String.drop_prefix s pos
;;

(* indices into the array hcj_ml_frame_info returns *)
let fi_ncomp = 2
let fi_decoded_w i = 7 + i
let fi_decoded_h i = 7 + 4 + i
let fi_actual_w i = 7 + 8 + i
let fi_actual_h i = 7 + 12 + i

module Header = struct
  type t =
    { frame : Markers.Sof.t option
    ; quant_tables : Markers.Dqt.t list
    ; huffman_tables : Markers.Dht.t list
    ; restart_interval : Markers.Dri.t option
    ; scan : Markers.Sos.t option
    }
  [@@deriving sexp_of]

  let of_raw (r : int array) =
    let k = ref 0 in
    let next () =
      let v = r.(!k) in
      Int.incr k;
      v
    in
    let has_frame = next () in
    let length = next () in
    let sample_precision = next () in
    let width = next () in
    let height = next () in
    let number_of_components = next () in
    let components =
      Array.init 4 ~f:(fun _ ->
          let identifier = next () in
          let horizontal_sampling_factor = next () in
          let vertical_sampling_factor = next () in
          let quantization_table_identifier = next () in
          { Markers.Component.identifier
          ; horizontal_sampling_factor
          ; vertical_sampling_factor
          ; quantization_table_identifier
          })
    in
    let frame =
      if has_frame = 0
      then None
      else
        Some
          { Markers.Sof.length
          ; sample_precision
          ; width
          ; height
          ; number_of_components
          ; components = Array.sub components ~pos:0 ~len:(min 4 number_of_components)
          }
    in
    let has_scan = next () in
    let sos_length = next () in
    let number_of_image_components = next () in
    let scan_components =
      Array.init 4 ~f:(fun _ ->
          let selector = next () in
          let dc_coef_selector = next () in
          let ac_coef_selector = next () in
          { Markers.Scan_component.selector; dc_coef_selector; ac_coef_selector })
    in
    let start_of_predictor_selection = next () in
    let end_of_predictor_selection = next () in
    let successive_approximation_bit_high = next () in
    let successive_approximation_bit_low = next () in
    let scan =
      if has_scan = 0
      then None
      else
        Some
          { Markers.Sos.length = sos_length
          ; number_of_image_components
          ; scan_components =
              Array.sub scan_components ~pos:0 ~len:(min 4 number_of_image_components)
          ; start_of_predictor_selection
          ; end_of_predictor_selection
          ; successive_approximation_bit_high
          ; successive_approximation_bit_low
          }
    in
    let has_restart_interval = next () in
    let dri_length = next () in
    let ri = next () in
    let restart_interval =
      if has_restart_interval = 0
      then None
      else Some { Markers.Dri.length = dri_length; restart_interval = ri }
    in
    let _scan_byte_pos = next () in
    let n_quant_tables = next () in
    let n_huffman_tables = next () in
    let quant_tables =
      List.init n_quant_tables ~f:(fun _ ->
          let length = next () in
          let element_precision = next () in
          let table_identifier = next () in
          let elements = Array.init 64 ~f:(fun _ -> next ()) in
          { Markers.Dqt.length; element_precision; table_identifier; elements })
    in
    let huffman_tables =
      List.init n_huffman_tables ~f:(fun _ ->
          let length = next () in
          let table_class = next () in
          let destination_identifier = next () in
          let lengths = Array.init 16 ~f:(fun _ -> next ()) in
          let nvalues = next () in
          let values = Array.init nvalues ~f:(fun _ -> next ()) in
          { Markers.Dht.length; table_class; destination_identifier; lengths; values })
    in
    { frame; quant_tables; huffman_tables; restart_interval; scan }
  ;;

  let decode bits = of_raw (header_decode_raw (file_of bits))
end

module Decoder = struct
  let frame_of_yuv info out =
    let aw i = info.(fi_actual_w i)
    and ah i = info.(fi_actual_h i) in
    let y = plane_of out ~pos:0 ~width:(aw 0) ~height:(ah 0) in
    let u = plane_of out ~pos:(aw 0 * ah 0) ~width:(aw 1) ~height:(ah 1) in
    let v = plane_of out ~pos:((aw 0 * ah 0) + (aw 1 * ah 1)) ~width:(aw 2) ~height:(ah 2) in
    Frame.of_planes ~y ~u ~v
  ;;

  let decode_frames bits =
    let files = Array.map bits ~f:file_of in
    let outs = decode_batch files 0 in
    Array.map2_exn files outs ~f:(fun s out -> frame_of_yuv (frame_info s) out)
  ;;

  let decode_a_frame bits = (decode_frames [| bits |]).(0)

  let decoded_planes bits =
    let s = file_of bits in
    let info = frame_info s in
    let out = (decode_batch [| s |] 1).(0) in
    let pos = ref 0 in
    Array.init info.(fi_ncomp) ~f:(fun i ->
        let width = info.(fi_decoded_w i)
        and height = info.(fi_decoded_h i) in
        let p = plane_of out ~pos:!pos ~width ~height in
        pos := !pos + (width * height);
        p)
  ;;

  let decode_rgb24 bits = (decode_batch [| file_of bits |] 2).(0)

  module Block = struct
    type t =
      { x : int
      ; y : int
      ; dc_pred : int
      ; identifier : int
      ; coefs : int array
      ; dequant : int array
      ; idct : int array
      ; recon : int array
      }

    (* the s-expression of Decoder.Component.Summary (decoder.ml:192-202), labels included *)
    let sexp_of_t { x; y; dc_pred; identifier; coefs; dequant; idct; recon } =
      let component =
        { Markers.Component.identifier
        ; horizontal_sampling_factor = 0
        ; vertical_sampling_factor = 0
        ; quantization_table_identifier = 0
        }
      in
      [%message
        (x : int)
          (y : int)
          (dc_pred : int)
          (component.identifier : int)
          (coefs : Util.coef_block)
          (dequant : Util.coef_block)
          (idct : Util.pixel_block)
          (recon : Util.pixel_block)]
    ;;
  end

  (* struct hcj_block_log (include/hcjpeg.h): 4 x int32, 64 x int16, 64 x int32, 64 x int32, 64 x uint8 = 720 bytes *)
  let block_log_bytes = 16 + 128 + 256 + 256 + 64

  let decode_log bits =
    let s = file_of bits in
    let header = Header.of_raw (header_decode_raw s) in
    let selectors =
      match header.scan with
      | Some scan -> Array.map scan.scan_components ~f:(fun c -> c.selector)
      | None -> [||]
    in
    let raw = decode_log_raw s in
    let n = Bigstring.length raw / block_log_bytes in
    Array.init n ~f:(fun b ->
        let base = b * block_log_bytes in
        let i32 pos = Bigstring.get_int32_le raw ~pos:(base + pos) in
        let i16 pos = Bigstring.get_int16_le raw ~pos:(base + pos) in
        let component = i32 12 in
        { Block.x = i32 0
        ; y = i32 4
        ; dc_pred = i32 8
        ; identifier =
            (if component < Array.length selectors then selectors.(component) else component)
        ; coefs = Array.init 64 ~f:(fun k -> i16 (16 + (2 * k)))
        ; dequant = Array.init 64 ~f:(fun k -> i32 (144 + (4 * k)))
        ; idct = Array.init 64 ~f:(fun k -> i32 (400 + (4 * k)))
        ; recon = Array.init 64 ~f:(fun k -> Bigstring.get_uint8 raw ~pos:(base + 656 + k))
        })
  ;;
end

module Encoder = struct
  let chroma_of frame =
    match Frame.chroma_subsampling frame with
    | Frame.Chroma_subsampling.C420 -> 420
    | C422 -> 422
    | C444 -> 444
  ;;

  (* planar Y,U,V exactly as Frame.output writes it (frame.ml:66-70) *)
  let bytes_of_frame frame =
    let planes = [ Frame.y frame; Frame.u frame; Frame.v frame ] in
    let total = List.sum (module Int) planes ~f:(fun p -> Plane.width p * Plane.height p) in
    let buf = Bigstring.create total in
    let (_ : int) = List.fold planes ~init:0 ~f:(fun pos p -> plane_into p buf ~pos) in
    buf
  ;;

  (* Append a finished file to the caller's writer: Bitstream_writer has no bulk entry point
     (bitstream_writer.mli:6), so byte by byte, without stuffing (the file is already stuffed). *)
  let append writer bytes =
    String.iter bytes ~f:(fun c ->
        Bitstream_writer.put_bits writer ~stuffing:false ~value:(Char.to_int c) ~bits:8)
  ;;

  let encode_frames_chroma ~frames ~quality ~chroma =
    if Array.is_empty frames
    then [||]
    else (
      let width = Frame.width frames.(0)
      and height = Frame.height frames.(0) in
      encode_batch (Array.map frames ~f:bytes_of_frame) width height chroma quality)
  ;;

  let encode_frames ~frames ~quality =
    if Array.is_empty frames
    then [||]
    else encode_frames_chroma ~frames ~quality ~chroma:(chroma_of frames.(0))
  ;;

  let encode_yuv ~frame ~quality ~writer ~chroma =
    append writer (encode_frames_chroma ~frames:[| frame |] ~quality ~chroma).(0)
  ;;

  let encode_420 ~frame ~quality ~writer = encode_yuv ~frame ~quality ~writer ~chroma:420
  let encode_422 ~frame ~quality ~writer = encode_yuv ~frame ~quality ~writer ~chroma:422
  let encode_444 ~frame ~quality ~writer = encode_yuv ~frame ~quality ~writer ~chroma:444

  (* Encoder.encode_monochrome (encoder.ml:543-552): one plane; chroma = 400 selects Parameters.monochrome *)
  let encode_monochrome ~frame ~quality ~writer =
    let buf = Bigstring.create (Plane.width frame * Plane.height frame) in
    let (_ : int) = plane_into frame buf ~pos:0 in
    append writer (encode_batch [| buf |] (Plane.width frame) (Plane.height frame) 400 quality).(0)
  ;;

  (* Entries are the model's own Encoder.Block.t (encoder.mli:19-43: a concrete record), so printing them with
     Encoder.Block.sexp_of_t gives the model's log verbatim.
     struct hcj_encoder_block (include/hcjpeg.h): 5 x int32, 64 x uint8, 64 x int32, 3 x 64 x int16, 2 x 64 x int32,
     2 x 64 x uint8 = 20 + 64 + 256 + 384 + 512 + 128 = 1364 bytes *)
  let encoder_block_bytes = 1364

  let encode_log ~frame ~quality =
    let raw =
      encode_log_raw (bytes_of_frame frame) (Frame.width frame) (Frame.height frame) (chroma_of frame) quality
    in
    let n = Bigstring.length raw / encoder_block_bytes in
    Array.init n ~f:(fun b ->
        let base = b * encoder_block_bytes in
        let i32 pos = Bigstring.get_int32_le raw ~pos:(base + pos) in
        let i16 pos = Bigstring.get_int16_le raw ~pos:(base + pos) in
        let u8 pos = Bigstring.get_uint8 raw ~pos:(base + pos) in
        let nrle = i32 16 in
        { Model_encoder.Block.x_pos = i32 0
        ; y_pos = i32 4
        ; dc_pred = i32 8
        ; input_pixels = Array.init 64 ~f:(fun k -> u8 (20 + k))
        ; fdct = Array.init 64 ~f:(fun k -> i32 (84 + (4 * k)))
        ; quant = Array.init 64 ~f:(fun k -> i16 (340 + (2 * k)))
        ; rle =
            List.init nrle ~f:(fun k ->
                { Model_encoder.Rle.run = i16 (468 + (2 * k)); value = i16 (596 + (2 * k)) })
        ; decoded =
            Some
              { Model_encoder.Block.Decoded.dequant = Array.init 64 ~f:(fun k -> i32 (724 + (4 * k)))
              ; idct = Array.init 64 ~f:(fun k -> i32 (980 + (4 * k)))
              ; recon = Array.init 64 ~f:(fun k -> u8 (1236 + k))
              ; error = Array.init 64 ~f:(fun k -> u8 (1300 + k))
              }
        })
  ;;
end
