// hcj_device.cuh — per-thread building blocks of the kernels, written as __host__ __device__ functions so
// that tests/emul can compile the very same code with g++ and check it against the oracle on the CPU
// before any GPU time is spent.  (The emulation is test infrastructure; the product only ever runs
// these functions inside the CUDA kernels of hcj_kernels.cu.)
#pragma once
#include <stdint.h>

#include "hcj_common.h"

#if defined(__CUDACC__)
#define HCJ_HD __host__ __device__ __forceinline__
#else
#define HCJ_HD inline
#endif

namespace hcjdev {

// ------------------------------------------------------------------------------------------------
// Bit reader over the destuffed entropy-coded bytes of one image.
//
// Semantics of Bitstream_reader.From_string (common/src/bitstream_reader.ml:19-38): MSB first, bits at
// or beyond `end_bits` read as zero.  `pos` is an absolute bit position in the image's entropy buffer.
// ------------------------------------------------------------------------------------------------
HCJ_HD uint32_t bswap32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(x, 0, 0x0123);
#else
  return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
#endif
}

HCJ_HD uint32_t load_be_word(const uint32_t *__restrict__ words, uint32_t idx, uint32_t end_bits) {
  uint32_t bitbase = idx << 5;
  if (bitbase >= end_bits) return 0u;
#if defined(__CUDA_ARCH__)
  uint32_t w = bswap32(__ldg(words + idx));
#else
  uint32_t w = bswap32(words[idx]);
#endif
  uint32_t rem = end_bits - bitbase;
  if (rem < 32u) w &= ~(0xffffffffu >> rem);
  return w;
}

struct BitReader {
  const uint32_t *words;
  uint32_t end_bits;
  uint32_t pos;
  uint32_t widx;
  uint32_t w0, w1;

  HCJ_HD void init(const uint32_t *words_, uint32_t pos_, uint32_t end_bits_) {
    words = words_;
    end_bits = end_bits_;
    pos = pos_;
    widx = pos_ >> 5;
    w0 = load_be_word(words, widx, end_bits);
    w1 = load_be_word(words, widx + 1, end_bits);
  }
  // The next 32 bits, MSB-aligned.
  HCJ_HD uint32_t window() const {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(w1, w0, pos);
#else
    uint32_t s = pos & 31u;
    return s ? (w0 << s) | (w1 >> (32u - s)) : w0;
#endif
  }
  // Advance by n <= 32 bits.
  HCJ_HD void skip(uint32_t n) {
    pos += n;
    uint32_t nidx = pos >> 5;
    if (nidx != widx) {
      widx = nidx;
      w0 = w1;
      w1 = load_be_word(words, nidx + 1, end_bits);
    }
  }
};

// Huffman tables of one scan component as the decode loops see them.
struct Tables {
  const uint16_t *dc_primary, *ac_primary;  // HCJ_LUT_SIZE entries each (shared memory in the kernels)
  const uint16_t *dc_full, *ac_full;        // 2^max_bits entries (global memory)
  uint32_t dc_max_bits, ac_max_bits;
};

// Tables.Lut lookup (decoder.ml:89-105): (length << 8) | data, 0 = None.
HCJ_HD uint32_t lut_lookup(const uint16_t *primary, const uint16_t *full, uint32_t max_bits, uint32_t win) {
  uint32_t e = primary[win >> (32 - HCJ_LUT_BITS)];
  if (e == 0u && max_bits > HCJ_LUT_BITS) e = full[win >> (32u - max_bits)];
  return e;
}

// Decoder.mag' (decoder.ml:73-79) for cat >= 1 on the cat bits that follow the code.
HCJ_HD int32_t extend(uint32_t bits, uint32_t cat) {
  int32_t v = (int32_t)bits;
  return (bits >> (cat - 1u)) ? v : v - (int32_t)(1u << cat) + 1;
}

// ------------------------------------------------------------------------------------------------
// Exact decode of one 8x8 block (Decoder.huffman_decode, decoder.ml:118-140) into a zero-initialised
// int16 block in zig-zag order; slot 0 receives the RESOLVED dc (pred + diff, decoder.ml:143).
// `seg_bits` is the length of the reader the model would be using (for the `show` bound, reader.ml:32).
// Returns 0 or the status of the exception the model raises.
// ------------------------------------------------------------------------------------------------
HCJ_HD int decode_block_exact(BitReader &br, const Tables &t, uint32_t seg_bits, int32_t &dc_pred,
                              int16_t *__restrict__ blk) {
  const bool careful = seg_bits <= 16u;
  if (careful && t.dc_max_bits >= seg_bits) return HCJ_DEV_BITS_OOB;
  uint32_t win = br.window();
  uint32_t e = lut_lookup(t.dc_primary, t.dc_full, t.dc_max_bits, win);
  if (e == 0u) return HCJ_DEV_NO_DC_CODE;
  uint32_t len = e >> 8, cat = e & 0xffu;
  int32_t diff = 0;
  if (cat) {
    if (careful && cat >= seg_bits) return HCJ_DEV_BITS_OOB;
    diff = extend((win << len) >> (32u - cat), cat);
  }
  br.skip(len + cat);
  dc_pred += diff;
  if (dc_pred < -32768 || dc_pred > 32767) return HCJ_DEV_DC_RANGE;
  blk[0] = (int16_t)dc_pred;
  uint32_t k = 1;
  while (k < 64u) {
    if (careful && t.ac_max_bits >= seg_bits) return HCJ_DEV_BITS_OOB;
    win = br.window();
    e = lut_lookup(t.ac_primary, t.ac_full, t.ac_max_bits, win);
    if (e == 0u) return HCJ_DEV_NO_AC_CODE;
    len = e >> 8;
    uint32_t rs = e & 0xffu, size = rs & 15u;
    if (careful && size >= seg_bits && size) return HCJ_DEV_BITS_OOB;
    uint32_t mbits = size ? (win << len) >> (32u - size) : 0u;
    br.skip(len + size);
    if (rs == 0u) break;  // (run, size) = (0, 0): end of block (decoder.ml:131-132)
    k += rs >> 4;
    if (k >= 64u) return HCJ_DEV_COEF_INDEX;
    if (size) blk[k] = (int16_t)extend(mbits, size);  // size 0 (e.g. ZRL) stores 0: already there
    k++;
  }
  return HCJ_DEV_OK;
}

// ------------------------------------------------------------------------------------------------
// Self-synchronising subsequence decode (no restart markers).
//
// The scan is cut into subsequences of S bits.  A decoder state between symbols is (p, c, z):
// bit position, block-in-MCU index, next zig-zag index (0 = a DC symbol comes next).  `sync` decodes
// the symbols that START inside [p, hi) without storing coefficients and reports the state at the
// first symbol start >= hi together with what the final pass needs for its prefix sums: the number of
// blocks begun and the sum of DC differentials per scan component.  On undefined codes / overlong runs
// (only possible while speculating from a wrong state, or in a corrupt stream) it recovers
// deterministically: skip one bit / close the block.  The final pass reports those as errors.
// ------------------------------------------------------------------------------------------------
struct ScanCtx {
  const uint32_t *words;       // destuffed bytes of the image (16-byte aligned)
  uint32_t total_bits;         // 8 * destuffed length
  uint32_t bpm;                // blocks per MCU
  const uint8_t *blk_comp;     // [bpm] block-in-MCU -> scan component
  Tables tab[HCJ_MAX_COMP];    // per scan component
};

struct SubResult {
  uint32_t p, cz;     // end state: cz = (c << 8) | z
  uint32_t nstart;    // DC symbols (blocks begun) decoded
  int32_t dcsum[HCJ_MAX_COMP];
};

HCJ_HD void subseq_sync(const ScanCtx &sc, uint32_t p, uint32_t cz, uint32_t hi, SubResult &r) {
  uint32_t c = cz >> 8, z = cz & 0xffu;
  uint32_t nstart = 0;
  int32_t d0 = 0, d1 = 0, d2 = 0, d3 = 0;
  BitReader br;
  br.init(sc.words, p, sc.total_bits);
  uint32_t comp = sc.blk_comp[c];
  while (br.pos < hi) {
    const Tables &t = sc.tab[comp];
    uint32_t win = br.window();
    if (z == 0u) {
      uint32_t e = lut_lookup(t.dc_primary, t.dc_full, t.dc_max_bits, win);
      if (e == 0u) {  // undefined code: resynchronise one bit later
        br.skip(1);
        continue;
      }
      uint32_t len = e >> 8, cat = e & 0xffu;
      int32_t diff = cat ? extend((win << len) >> (32u - cat), cat) : 0;
      br.skip(len + cat);
      d0 += comp == 0u ? diff : 0;
      d1 += comp == 1u ? diff : 0;
      d2 += comp == 2u ? diff : 0;
      d3 += comp == 3u ? diff : 0;
      nstart++;
      z = 1;
    } else {
      uint32_t e = lut_lookup(t.ac_primary, t.ac_full, t.ac_max_bits, win);
      if (e == 0u) {
        br.skip(1);
        continue;
      }
      uint32_t rs = e & 0xffu;
      br.skip((e >> 8) + (rs & 15u));
      z = rs ? z + (rs >> 4) + 1u : 64u;
      if (z >= 64u) {  // block complete (EOB, coefficient 63, or an overlong run while speculating)
        z = 0;
        c = c + 1u == sc.bpm ? 0u : c + 1u;
        comp = sc.blk_comp[c];
      }
    }
  }
  r.p = br.pos;
  r.cz = (c << 8) | z;
  r.nstart = nstart;
  r.dcsum[0] = d0;
  r.dcsum[1] = d1;
  r.dcsum[2] = d2;
  r.dcsum[3] = d3;
}

// Final pass over one subsequence from its exact start state.  `blk` is the index (within the image)
// of the block in progress (start of a block: the index of the previous one), `pred` the DC predictors
// at the start state.  Stores coefficients (zig-zag, DC resolved) into the zero-initialised `coefs`.
// `hi` = end of the subsequence; the last subsequence passes 0xffffffff and runs until `nblocks` blocks
// are complete, reading zero bits past the end exactly as the model's reader does.
HCJ_HD int subseq_write(const ScanCtx &sc, uint32_t p, uint32_t cz, uint32_t hi, int64_t blk, int32_t pred[HCJ_MAX_COMP],
                        int64_t nblocks, int16_t *__restrict__ coefs, uint32_t *err_pos) {
  uint32_t c = cz >> 8, z = cz & 0xffu;
  BitReader br;
  br.init(sc.words, p, sc.total_bits);
  uint32_t comp = sc.blk_comp[c];
  int32_t p0 = pred[0], p1 = pred[1], p2 = pred[2], p3 = pred[3];
  if (z != 0u && blk >= nblocks) return HCJ_DEV_OK;
  int16_t *out = coefs + blk * 64;
  while (br.pos < hi) {
    const Tables &t = sc.tab[comp];
    uint32_t win = br.window();
    if (z == 0u) {
      if (blk + 1 >= nblocks) break;  // every block of the frame is complete
      uint32_t e = lut_lookup(t.dc_primary, t.dc_full, t.dc_max_bits, win);
      if (e == 0u) return *err_pos = br.pos, HCJ_DEV_NO_DC_CODE;
      uint32_t len = e >> 8, cat = e & 0xffu;
      int32_t diff = cat ? extend((win << len) >> (32u - cat), cat) : 0;
      br.skip(len + cat);
      int32_t v;
      if (comp == 0u) v = (p0 += diff);
      else if (comp == 1u) v = (p1 += diff);
      else if (comp == 2u) v = (p2 += diff);
      else v = (p3 += diff);
      if (v < -32768 || v > 32767) return *err_pos = br.pos, HCJ_DEV_DC_RANGE;
      blk++;
      out = coefs + blk * 64;
      out[0] = (int16_t)v;
      z = 1;
    } else {
      uint32_t e = lut_lookup(t.ac_primary, t.ac_full, t.ac_max_bits, win);
      if (e == 0u) return *err_pos = br.pos, HCJ_DEV_NO_AC_CODE;
      uint32_t len = e >> 8, rs = e & 0xffu, size = rs & 15u;
      uint32_t mbits = size ? (win << len) >> (32u - size) : 0u;
      br.skip(len + size);
      if (rs == 0u) {
        z = 64;
      } else {
        z += rs >> 4;
        if (z >= 64u) return *err_pos = br.pos, HCJ_DEV_COEF_INDEX;
        if (size) out[z] = (int16_t)extend(mbits, size);
        z++;
      }
      if (z >= 64u) {
        z = 0;
        c = c + 1u == sc.bpm ? 0u : c + 1u;
        comp = sc.blk_comp[c];
      }
    }
  }
  return HCJ_DEV_OK;
}

// ------------------------------------------------------------------------------------------------
// Dct.Chen.inverse_8x8 (dct.ml:11-107), bit-exact.
//
// T = int32_t: every value is held in 32 bits except the two `181 *` products per pass (dct.ml:43-44,
// 87-88), which are formed in 64 bits.  That is exact whenever sum(|dequantised coefficient|) over the
// block is below HCJ_IDCT_L1_LIMIT (derivation in DESIGN.md; checked in tests/test_emul_idct.py).
// T = int64_t is the model's arithmetic verbatim and is taken for the (pathological) blocks above.
// ------------------------------------------------------------------------------------------------
#define HCJ_IDCT_L1_LIMIT 60000

template <typename T>
HCJ_HD T mul181(T a) {
  return (T)(((int64_t)a * 181 + 128) >> 8);
}

template <typename T, int STRIDE>
HCJ_HD void idct_row(T *b) {  // dct.ml:11-54
  const T W1 = 2841, W2 = 2676, W3 = 2408, W5 = 1609, W6 = 1108, W7 = 565;
  T x0 = b[0 * STRIDE] * 2048 + 128, x1 = b[4 * STRIDE] * 2048, x2 = b[6 * STRIDE], x3 = b[2 * STRIDE],
    x4 = b[1 * STRIDE], x5 = b[7 * STRIDE], x6 = b[5 * STRIDE], x7 = b[3 * STRIDE], x8;
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
  x2 = mul181<T>(x4 + x5);
  x4 = mul181<T>(x4 - x5);
  b[0 * STRIDE] = (x7 + x1) >> 8;
  b[1 * STRIDE] = (x3 + x2) >> 8;
  b[2 * STRIDE] = (x0 + x4) >> 8;
  b[3 * STRIDE] = (x8 + x6) >> 8;
  b[4 * STRIDE] = (x8 - x6) >> 8;
  b[5 * STRIDE] = (x0 - x4) >> 8;
  b[6 * STRIDE] = (x3 - x2) >> 8;
  b[7 * STRIDE] = (x7 - x1) >> 8;
}

template <typename T, int STRIDE>
HCJ_HD void idct_col(T *b) {  // dct.ml:56-98
  const T W1 = 2841, W2 = 2676, W3 = 2408, W5 = 1609, W6 = 1108, W7 = 565;
  T x0 = b[0 * STRIDE] * 256 + 8192, x1 = b[4 * STRIDE] * 256, x2 = b[6 * STRIDE], x3 = b[2 * STRIDE],
    x4 = b[1 * STRIDE], x5 = b[7 * STRIDE], x6 = b[5 * STRIDE], x7 = b[3 * STRIDE], x8;
  x8 = W7 * (x4 + x5) + 4;
  x4 = (x8 + (W1 - W7) * x4) >> 3;
  x5 = (x8 - (W1 + W7) * x5) >> 3;
  x8 = W3 * (x6 + x7) + 4;
  x6 = (x8 - (W3 - W5) * x6) >> 3;
  x7 = (x8 - (W3 + W5) * x7) >> 3;
  x8 = x0 + x1;
  x0 = x0 - x1;
  x1 = W6 * (x3 + x2) + 4;
  x2 = (x1 - (W2 + W6) * x2) >> 3;
  x3 = (x1 + (W2 - W6) * x3) >> 3;
  x1 = x4 + x6;
  x4 = x4 - x6;
  x6 = x5 + x7;
  x5 = x5 - x7;
  x7 = x8 + x3;
  x8 = x8 - x3;
  x3 = x0 + x2;
  x0 = x0 - x2;
  x2 = mul181<T>(x4 + x5);
  x4 = mul181<T>(x4 - x5);
  b[0 * STRIDE] = (x7 + x1) >> 14;
  b[1 * STRIDE] = (x3 + x2) >> 14;
  b[2 * STRIDE] = (x0 + x4) >> 14;
  b[3 * STRIDE] = (x8 + x6) >> 14;
  b[4 * STRIDE] = (x8 - x6) >> 14;
  b[5 * STRIDE] = (x0 - x4) >> 14;
  b[6 * STRIDE] = (x3 - x2) >> 14;
  b[7 * STRIDE] = (x7 - x1) >> 14;
}

template <typename T>
HCJ_HD void idct_8x8(T v[64]) {  // dct.ml:100-107: all rows, then all columns
#pragma unroll
  for (int i = 0; i < 8; i++) idct_row<T, 1>(v + 8 * i);
#pragma unroll
  for (int i = 0; i < 8; i++) idct_col<T, 8>(v + i);
}

// Zigzag.inverse (zigzag.ml:3-69) as a compile-time function so that fully unrolled loops index
// registers, not memory.
HCJ_HD constexpr int zigzag_inverse(int i) {
  constexpr int t[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                         41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                         30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  return t[i];
}
// Zigzag.forward (zigzag.ml:71-137): natural index -> zig-zag position.
HCJ_HD constexpr int zigzag_forward(int i) {
  constexpr int t[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30,
                         41, 43, 9,  11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38,
                         46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};
  return t[i];
}

// dequantize + inverse zig-zag (decoder.ml:142-149), IDCT, clip and level shift (decoder.ml:213-224).
// c[i]: zig-zag coefficients with c[0] the resolved DC; q[i]: quant table in file (zig-zag) order.
// pix[k]: reconstructed sample k = x + 8*y.  `force_wide`: quant entries above 255 are present.
HCJ_HD void reconstruct_block(const int16_t c[64], const uint16_t q[64], bool force_wide, uint8_t pix[64]) {
  int32_t v[64];
  uint32_t l1 = 0;
  bool wide = force_wide;
  if (!force_wide) {
#pragma unroll
    for (int i = 0; i < 64; i++) {
      int32_t d = (int32_t)c[i] * (int32_t)q[i];
      v[zigzag_inverse(i)] = d;
      l1 += (uint32_t)(d < 0 ? -d : d);
    }
    wide = l1 >= (uint32_t)HCJ_IDCT_L1_LIMIT;
  }
  if (!wide) {
    idct_8x8<int32_t>(v);
#pragma unroll
    for (int k = 0; k < 64; k++) {
      int32_t s = v[k] < -128 ? -128 : v[k] > 127 ? 127 : v[k];
      pix[k] = (uint8_t)(s + 128);
    }
  } else {
    int64_t w[64];
    for (int i = 0; i < 64; i++) w[zigzag_inverse(i)] = (int64_t)c[i] * (int64_t)q[i];
    idct_8x8<int64_t>(w);
    for (int k = 0; k < 64; k++) {
      int64_t s = w[k] < -128 ? -128 : w[k] > 127 ? 127 : w[k];
      pix[k] = (uint8_t)(s + 128);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Dct.Chen.forward_8x8 (dct.ml:109-196): columns first, then rows.  Inputs are pixel - 128, so 32-bit
// arithmetic is exact (|value| < 2^21 throughout).
// ------------------------------------------------------------------------------------------------
HCJ_HD int32_t fc4(int32_t f, int32_t g) { return (362 * (f + g)) >> 9; }
HCJ_HD int32_t fc62(int32_t f, int32_t g) { return (196 * f + 473 * g) >> 9; }
HCJ_HD int32_t fc71(int32_t f, int32_t g) { return (100 * f + 502 * g) >> 9; }
HCJ_HD int32_t fc35(int32_t f, int32_t g) { return (426 * f + 284 * g) >> 9; }

template <int STRIDE>
HCJ_HD void fdct_1d(int32_t *b) {  // dct.ml:114-149 / :151-187
  int32_t a0 = b[0 * STRIDE] + b[7 * STRIDE], c3 = b[0 * STRIDE] - b[7 * STRIDE];
  int32_t a1 = b[1 * STRIDE] + b[6 * STRIDE], c2 = b[1 * STRIDE] - b[6 * STRIDE];
  int32_t a2 = b[2 * STRIDE] + b[5 * STRIDE], c1 = b[2 * STRIDE] - b[5 * STRIDE];
  int32_t a3 = b[3 * STRIDE] + b[4 * STRIDE], c0 = b[3 * STRIDE] - b[4 * STRIDE];
  int32_t b0 = a0 + a3, b1 = a1 + a2, b2 = a1 - a2, b3 = a0 - a3;
  b[0 * STRIDE] = fc4(b0, b1);
  b[4 * STRIDE] = fc4(b0, -b1);
  b[2 * STRIDE] = fc62(b2, b3);
  b[6 * STRIDE] = fc62(b3, -b2);
  b0 = fc4(c2, -c1);
  b1 = fc4(c2, c1);
  a0 = c0 + b0;
  a1 = c0 - b0;
  a2 = c3 - b1;
  a3 = c3 + b1;
  b[1 * STRIDE] = fc71(a0, a3);
  b[5 * STRIDE] = fc35(a1, a2);
  b[3 * STRIDE] = fc35(a2, -a1);
  b[7 * STRIDE] = fc71(a3, -a0);
}

HCJ_HD void fdct_8x8(int32_t v[64]) {
#pragma unroll
  for (int i = 0; i < 8; i++) fdct_1d<8>(v + i);
#pragma unroll
  for (int i = 0; i < 8; i++) fdct_1d<1>(v + 8 * i);
}

// Encoder.quant_and_scale (encoder.ml:98-101): (f -/+ 2q) / (4q), OCaml's truncating division,
// computed with a reciprocal: recip = floor(2^32 / (4q)) + 1 is exact for |f| + 2q < 2^20
// (checked exhaustively in tests/test_emul_encode.py).
HCJ_HD int32_t quantize(int32_t f, uint32_t q, uint32_t recip) {
  uint32_t n = (uint32_t)(f < 0 ? -f : f) + 2u * q;
#if defined(__CUDA_ARCH__)
  uint32_t d = __umulhi(n, recip);
#else
  uint32_t d = (uint32_t)(((uint64_t)n * recip) >> 32);
#endif
  return f < 0 ? -(int32_t)d : (int32_t)d;
}

// Encoder.size (encoder.ml:143) and Encoder.magnitude (encoder.ml:145-147).
HCJ_HD uint32_t coef_size(int32_t v) {
  uint32_t a = (uint32_t)(v < 0 ? -v : v);
#if defined(__CUDA_ARCH__)
  return 32u - (uint32_t)__clz((int)a);
#else
  uint32_t s = 0;
  while (a) {
    s++;
    a >>= 1;
  }
  return s;
#endif
}
HCJ_HD uint32_t coef_magnitude(int32_t v, uint32_t size) {
  return (uint32_t)(v >= 0 ? v : v - 1) & ((1u << size) - 1u);
}

}  // namespace hcjdev

namespace hcjdev {

// ------------------------------------------------------------------------------------------------
// Encoder.rle + write_bits (encoder.ml:127-193) for one block: calls emit(bits, nbits) for every
// field the model passes to Writer.put_bits, in order.  q: quantised zig-zag block (q[0] ignored),
// dcdiff: quant.(0) - dc_pred.  dc_codes[size], ac_codes[(run << 4) | size] = (code << 8) | length.
// Returns false if a symbol has no code (the model raises an index-out-of-bounds exception).
// ------------------------------------------------------------------------------------------------
template <class Emit>
HCJ_HD bool encode_block_fields(const int16_t *q, int32_t dcdiff, const uint32_t *dc_codes, const uint32_t *ac_codes,
                                Emit &emit) {
  bool ok = true;
  uint32_t size = coef_size(dcdiff);
  uint32_t code = size < 16u ? dc_codes[size] : 0u;
  ok &= (code & 0xffu) != 0u;
  emit(code >> 8, code & 0xffu);
  emit(coef_magnitude(dcdiff, size), size);
  uint32_t run = 0;
  for (int k = 1; k < 64; k++) {
    int32_t v = q[k];
    if (k == 63 && v == 0) {  // [ { run; value = 0 } ] -> end of block (encoder.ml:172-175)
      code = ac_codes[0x00];
      ok &= (code & 0xffu) != 0u;
      emit(code >> 8, code & 0xffu);
    } else if (v != 0) {
      while (run >= 16u) {  // runs (encoder.ml:178-185)
        code = ac_codes[0xf0];
        ok &= (code & 0xffu) != 0u;
        emit(code >> 8, code & 0xffu);
        run -= 16u;
      }
      size = coef_size(v);
      code = size < 16u ? ac_codes[(run << 4) | size] : 0u;
      ok &= (code & 0xffu) != 0u;
      emit(code >> 8, code & 0xffu);
      emit(coef_magnitude(v, size), size);
      run = 0;
    } else {
      run++;
    }
  }
  return ok;
}

}  // namespace hcjdev
