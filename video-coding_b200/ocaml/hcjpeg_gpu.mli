(** GPU drop-in for the entry points of [Hardcaml_jpeg_model.Decoder] / [.Encoder] that [jpeg/bin/model.ml] and the
    test benches use.  Same types, same exceptions-as-messages; results are bit-identical to the software model (see the
    parity tests of the B200 repository).  Every identifier of the reference used here is exported by its [.mli]
    (table in INTEGRATION.md): [Plane.t] is abstract ([common/src/plane.mli:6]), so planes are filled and read through
    [Plane.create] and the [.!()] accessors only. *)

open Hardcaml_video_common

(** Number of CUDA devices visible to the process; batches are spread over all of them (capped by the environment
    variable [HCJPEG_GPUS]). *)
val device_count : unit -> int

module Header : sig
  (** Same fields, in the same order, as [Decoder.Header.t] (decoder.ml:6-13), which is abstract in decoder.mli:9 and
      cannot be constructed outside the model; the marker records are the model's own ([Markers], markers.mli). *)
  type t =
    { frame : Hardcaml_jpeg_model.Markers.Sof.t option
    ; quant_tables : Hardcaml_jpeg_model.Markers.Dqt.t list
    ; huffman_tables : Hardcaml_jpeg_model.Markers.Dht.t list
    ; restart_interval : Hardcaml_jpeg_model.Markers.Dri.t option
    ; scan : Hardcaml_jpeg_model.Markers.Sos.t option
    }
  [@@deriving sexp_of]

  (** [Decoder.Header.decode] (decoder.mli:13, decoder.ml:37-70) through [hcj_header_decode]. *)
  val decode : Hardcaml_jpeg_model.Decoder.Bits.t -> t
end

module Decoder : sig
  (** [Decoder.decode_a_frame] (decoder.mli:59, decoder.ml:422-427). *)
  val decode_a_frame : Hardcaml_jpeg_model.Decoder.Bits.t -> Frame.t

  (** The same for many files at once: one call, all GPUs (images sharded by index). *)
  val decode_frames : Hardcaml_jpeg_model.Decoder.Bits.t array -> Frame.t array

  (** [Decoder.get_decoded_planes] after [init] + [decode] (decoder.mli:49, decoder.ml:399-401): padded planes in
      scan order. *)
  val decoded_planes : Hardcaml_jpeg_model.Decoder.Bits.t -> Plane.t array

  (** Extension: interleaved RGB24 ([width * height * 3] bytes) after [Planar_444] up-sampling. *)
  val decode_rgb24 : Hardcaml_jpeg_model.Decoder.Bits.t -> Base_bigstring.t

  (** One entry per block of [Decoder.For_testing.Sequenced.decode] (decoder.mli:83), with the fields that
      [Decoder.Component.Summary] prints (decoder.ml:189-203) and the same s-expression. *)
  module Block : sig
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
    [@@deriving sexp_of]
  end

  (** [model decode log] (jpeg/bin/model.ml:46-68). *)
  val decode_log : Hardcaml_jpeg_model.Decoder.Bits.t -> Block.t array
end

module Encoder : sig
  (** [Encoder.encode_420 / 422 / 444 / monochrome] (encoder.mli:132-135, encoder.ml:522-552). *)
  val encode_420 : frame:Frame.t -> quality:int -> writer:Bitstream_writer.t -> unit

  val encode_422 : frame:Frame.t -> quality:int -> writer:Bitstream_writer.t -> unit
  val encode_444 : frame:Frame.t -> quality:int -> writer:Bitstream_writer.t -> unit
  val encode_monochrome : frame:Plane.t -> quality:int -> writer:Bitstream_writer.t -> unit

  (** Many frames of one geometry at once, all GPUs; the files in frame order. *)
  val encode_frames : frames:Frame.t array -> quality:int -> string array

  (** [model encode log -verbose] (jpeg/bin/model.ml:108-142): one entry per block of [Encoder.encode_seq]
      (encoder.mli:127), as the model's own [Encoder.Block.t] (encoder.mli:19-43, a concrete record) with [decoded]
      always present: [Encoder.Block.sexp_of_t] prints the model's log verbatim. *)
  val encode_log : frame:Frame.t -> quality:int -> Hardcaml_jpeg_model.Encoder.Block.t array
end
