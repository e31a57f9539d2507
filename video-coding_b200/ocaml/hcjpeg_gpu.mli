(** GPU drop-in for the entry points of [Hardcaml_jpeg_model.Decoder] / [.Encoder] that
    [jpeg/bin/model.ml] uses (decode frame / encode frame).  Same types, same exceptions-as-messages;
    results are bit-identical to the software model (see the parity tests of the B200 repository). *)

open Hardcaml_video_common

module Decoder : sig
  (** [Hardcaml_jpeg_model.Decoder.decode_a_frame] (decoder.ml:422-427). *)
  val decode_a_frame : Hardcaml_jpeg_model.Decoder.Bits.t -> Frame.t

  (** [Decoder.get_decoded_planes] after [init] + [decode] (decoder.ml:399-401): padded planes. *)
  val decoded_planes : Hardcaml_jpeg_model.Decoder.Bits.t -> Plane.t array
end

module Encoder : sig
  (** [Hardcaml_jpeg_model.Encoder.encode_420 / 422 / 444] (encoder.ml:522-541). *)
  val encode_420 : frame:Frame.t -> quality:int -> writer:Bitstream_writer.t -> unit

  val encode_422 : frame:Frame.t -> quality:int -> writer:Bitstream_writer.t -> unit
  val encode_444 : frame:Frame.t -> quality:int -> writer:Bitstream_writer.t -> unit
end
