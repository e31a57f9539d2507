(* model_gpu.ml - the GPU twin of jpeg/bin/model.ml: the same command tree (decode frame / header / log, encode
   frame / log), the same arguments and the same output formats, with the model's work done by libhcjpeg through
   Hardcaml_jpeg_gpu.  Not compiled in this repository (no OCaml toolchain in its image); every identifier of the
   reference it uses is listed with the .mli that exports it in INTEGRATION.md. *)
open Core

include struct
  open Hardcaml_jpeg_model
  module Decoder = Decoder
  module Encoder = Encoder
end

include struct
  open Hardcaml_video_common
  module Frame = Frame
  module Writer = Bitstream_writer
  module Size = Size
end

module Gpu = Hardcaml_jpeg_gpu.Hcjpeg_gpu

(* jpeg/bin/model.ml:18-27 *)
let command_decode_header =
  Command.basic
    ~summary:"Decode a frame header"
    [%map_open.Command
      let bits = anon ("INPUT-BITS" %: string) in
      fun () ->
        let bits = Decoder.Bits.create (In_channel.read_all bits) in
        let header = Gpu.Header.decode bits in
        print_s [%message (header : Gpu.Header.t)]]
;;

(* jpeg/bin/model.ml:29-44 *)
let command_decode_frame =
  Command.basic
    ~summary:"Decode a frame to YUV"
    [%map_open.Command
      let bits = anon ("INPUT-BITS" %: string)
      and yuv = anon (maybe ("OUTPUT-FRAME" %: string)) in
      fun () ->
        let bits = Decoder.Bits.create (In_channel.read_all bits) in
        let frame = Gpu.Decoder.decode_a_frame bits in
        let yuv =
          match yuv with
          | None -> Out_channel.stdout
          | Some yuv -> Out_channel.create yuv
        in
        Frame.output frame yuv]
;;

(* jpeg/bin/model.ml:46-68 *)
let command_decode_log =
  Command.basic
    ~summary:"Decode a frame and write a log file"
    [%map_open.Command
      let bits = anon ("INPUT-BITS" %: string) in
      fun () ->
        let bits = Decoder.Bits.create (In_channel.read_all bits) in
        let header = Gpu.Header.decode bits in
        print_s [%message (header : Gpu.Header.t)];
        Array.iteri (Gpu.Decoder.decode_log bits) ~f:(fun block_number component ->
            let block_number = ref block_number in
            print_s [%message (!block_number : int) (component : Gpu.Decoder.Block.t)])]
;;

(* jpeg/bin/model.ml:70-82 *)
let input_yuv file ~chroma_subsampling ~width ~height =
  let frame = Frame.create ~chroma_subsampling ~width ~height in
  In_channel.with_file file ~f:(fun file -> Frame.input frame file);
  frame
;;

let chroma_arg =
  Command.Arg_type.create (function
      | "420" -> Frame.Chroma_subsampling.C420
      | "422" -> Frame.Chroma_subsampling.C422
      | "444" -> Frame.Chroma_subsampling.C444
      | _ -> raise_s [%message "Invalid chroma type"])
;;

(* jpeg/bin/model.ml:84-106 *)
let command_encode_frame =
  Command.basic
    ~summary:"Encoder a frame"
    [%map_open.Command
      let yuv = anon ("INPUT-FRAME" %: string)
      and { width; height } = anon ("WIDTHxHEIGHT" %: Size.arg_type)
      and bits = anon ("OUTPUT-BITS" %: string)
      and quality = flag "-quality" (optional_with_default 75 int) ~doc:" Image quality"
      and chroma_subsampling =
        flag
          "-chroma"
          (optional_with_default Frame.Chroma_subsampling.C420 chroma_arg)
          ~doc:""
      in
      fun () ->
        let frame = input_yuv yuv ~chroma_subsampling ~width ~height in
        let writer = Writer.create () in
        (match chroma_subsampling with
        | C420 -> Gpu.Encoder.encode_420 ~frame ~quality ~writer
        | C422 -> Gpu.Encoder.encode_422 ~frame ~quality ~writer
        | C444 -> Gpu.Encoder.encode_444 ~frame ~quality ~writer);
        Out_channel.write_all bits ~data:(Writer.get_buffer writer)]
;;

(* jpeg/bin/model.ml:108-142; the reconstruction is always computed on the device, so -verbose only selects
   whether it is printed *)
let command_encode_log =
  Command.basic
    ~summary:"Encode a frame and write a log file"
    [%map_open.Command
      let yuv = anon ("INPUT-FRAME" %: string)
      and { width; height } = anon ("WIDTHxHEIGHT" %: Size.arg_type)
      and quality = flag "-quality" (optional_with_default 75 int) ~doc:" Image quality"
      and chroma_subsampling =
        flag
          "-chroma"
          (optional_with_default Frame.Chroma_subsampling.C420 chroma_arg)
          ~doc:""
      and verbose = flag "-verbose" no_arg ~doc:" Reconstruct and compute error" in
      fun () ->
        let frame = input_yuv yuv ~chroma_subsampling ~width ~height in
        Array.iteri (Gpu.Encoder.encode_log ~frame ~quality) ~f:(fun block_number block ->
            let block_number = ref block_number in
            let block : Encoder.Block.t =
              if verbose then block else { block with decoded = None }
            in
            print_s [%message (!block_number : int) (block : Encoder.Block.t)])]
;;

let () =
  Command_unix.run
    (Command.group
       ~summary:"JPEG encoder and decoder models (B200)"
       [ ( "decode"
         , Command.group
             ~summary:"Decoder model"
             [ "frame", command_decode_frame
             ; "header", command_decode_header
             ; "log", command_decode_log
             ] )
       ; ( "encode"
         , Command.group
             ~summary:"Encoder model"
             [ "frame", command_encode_frame; "log", command_encode_log ] )
       ])
;;
