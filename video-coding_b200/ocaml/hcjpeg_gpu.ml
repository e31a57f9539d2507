open Base
open Hardcaml_video_common
module Bigstring = Base_bigstring

external frame_info : string -> int array = "hcj_ml_frame_info"
external decode : string -> int -> Bigstring.t = "hcj_ml_decode"

external encode
  :  Bigstring.t
  -> int
  -> int
  -> int
  -> int
  -> string
  = "hcj_ml_encode"

(* Wrap [len] bytes of [buf] at [pos] as a Plane.t without changing the model's Plane representation. *)
let plane_of buf ~pos ~width ~height =
  let p = Plane.create ~width ~height in
  Bigstring.blit ~src:buf ~src_pos:pos ~dst:(Plane.plane p) ~dst_pos:0 ~len:(width * height);
  p
;;

module Decoder = struct
  let bits_to_string bits = Hardcaml_jpeg_model.Decoder.Bits.get_buffer bits

  let decode_a_frame bits =
    let s = bits_to_string bits in
    let info = frame_info s in
    let aw i = info.(7 + 8 + i)
    and ah i = info.(7 + 12 + i) in
    let out = decode s 0 in
    let y = plane_of out ~pos:0 ~width:(aw 0) ~height:(ah 0) in
    let u = plane_of out ~pos:(aw 0 * ah 0) ~width:(aw 1) ~height:(ah 1) in
    let v = plane_of out ~pos:((aw 0 * ah 0) + (aw 1 * ah 1)) ~width:(aw 2) ~height:(ah 2) in
    Frame.of_planes ~y ~u ~v
  ;;

  let decoded_planes bits =
    let s = bits_to_string bits in
    let info = frame_info s in
    let out = decode s 1 in
    let pos = ref 0 in
    Array.init info.(2) ~f:(fun i ->
        let width = info.(7 + i)
        and height = info.(7 + 4 + i) in
        let p = plane_of out ~pos:!pos ~width ~height in
        pos := !pos + (width * height);
        p)
  ;;
end

module Encoder = struct
  let encode_yuv ~frame ~quality ~writer ~chroma =
    let width = Frame.width frame
    and height = Frame.height frame in
    let planes = [ Frame.y frame; Frame.u frame; Frame.v frame ] in
    let total = List.sum (module Int) planes ~f:(fun p -> Plane.width p * Plane.height p) in
    let buf = Bigstring.create total in
    let _ =
      List.fold planes ~init:0 ~f:(fun pos p ->
          let len = Plane.width p * Plane.height p in
          Bigstring.blit ~src:(Plane.plane p) ~src_pos:0 ~dst:buf ~dst_pos:pos ~len;
          pos + len)
    in
    let bytes = encode buf width height chroma quality in
    (* Append to the caller's writer byte by byte: Bitstream_writer has no bulk entry point. *)
    String.iter bytes ~f:(fun c ->
        Bitstream_writer.put_bits writer ~stuffing:false ~bits:8 ~value:(Char.to_int c))
  ;;

  let encode_420 ~frame ~quality ~writer = encode_yuv ~frame ~quality ~writer ~chroma:420
  let encode_422 ~frame ~quality ~writer = encode_yuv ~frame ~quality ~writer ~chroma:422
  let encode_444 ~frame ~quality ~writer = encode_yuv ~frame ~quality ~writer ~chroma:444
end
