// emul.cpp — TEST INFRASTRUCTURE: compiles the kernels' per-thread building blocks (hcj_device.cuh)
// and the host half (hcj_host.cpp) with g++ and drives them with plain loops that stand in for the CUDA
// thread grid.  It lets `pytest -m "not gpu"` check the device logic against the oracle without a GPU.
// Nothing in the product links or calls this file.
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "hcj_device.cuh"
#include "hcj_host.h"

using namespace hcjdev;

namespace {

struct HostTables {
  std::vector<hcj::HuffLut> luts;  // [comp][dc/ac]
  // what the kernels keep in shared memory (table index = comp * 2 + (0 = dc, 1 = ac))
  std::vector<uint32_t> fast;
  std::vector<uint32_t> sub;
  std::vector<uint32_t> multi;  // multi-symbol AC entries of the synchronisation passes, [comp][HCJ_LUT_SIZE]
  uint32_t max_bits[HCJ_MAX_COMP * 2];
  const uint16_t *full[HCJ_MAX_COMP * 2];
  BlkInfo blkinfo[HCJ_MAX_BPM + 2];
  Tables tab[HCJ_MAX_COMP];
  int32_t quant[HCJ_MAX_COMP * 128];
  uint8_t blk_comp[HCJ_MAX_BPM + 2];
  FastTables fast_tables() const { return FastTables{fast.data(), sub.data(), max_bits, full, blkinfo, quant, multi.data()}; }
  Local local() const { return Local{fast_tables(), quant, blk_comp}; }
};

int prepare(const uint8_t *jpeg, size_t len, unsigned flags, hcj_header *h, hcj::ImagePlan *plan, HostTables *ht) {
  int st = hcj::header_decode(jpeg, len, h);
  if (st) return st;
  st = hcj::plan_image(*h, flags, plan);
  if (st) return st;
  ht->luts.resize(plan->info.ncomp * 2);
  for (int c = 0; c < plan->info.ncomp; c++) {
    st = hcj::build_lut(h->huffman_tables[plan->dc_index[c]], &ht->luts[c * 2]);
    if (st) return st;
    st = hcj::build_lut(h->huffman_tables[plan->ac_index[c]], &ht->luts[c * 2 + 1]);
    if (st) return st;
  }
  for (int c = 0; c < plan->info.ncomp; c++)
    for (int e = 0; e < 64; e++) {
      ht->quant[c * 128 + e] = h->quant_tables[plan->qt_index[c]].elements[e];
      ht->quant[c * 128 + 64 + e] = HCJ_QD(e, h->quant_tables[plan->qt_index[c]].elements[e] & 0xff);
    }
  ht->fast.assign((size_t)plan->info.ncomp * 2 * HCJ_LUT_SIZE, 0);
  ht->sub.assign((size_t)plan->info.ncomp * 2 * HCJ_LUT_NSUB * HCJ_LUT_SUB_SIZE, 0);
  for (int c = 0; c < plan->info.ncomp; c++) {
    for (int k = 0; k < 2; k++) {  // the kernels' table loader (load_tables in hcj_kernels.cu)
      const hcj::HuffLut &l = ht->luts[c * 2 + k];
      const int ti = c * 2 + k;
      ht->max_bits[ti] = (uint32_t)l.max_bits;
      ht->full[ti] = l.full.data();
      for (int i = 0; i < HCJ_LUT_SIZE; i++)
        ht->fast[(size_t)ti * HCJ_LUT_SIZE + i] = fast_entry_from_primary(l.primary[i], (uint32_t)l.max_bits, k == 0);
      for (int i = 0; i < HCJ_LUT_NSUB * HCJ_LUT_SUB_SIZE; i++)
        ht->sub[(size_t)ti * HCJ_LUT_NSUB * HCJ_LUT_SUB_SIZE + i] = fast_entry_or_none(
            l.primary[HCJ_LUT_SIZE + ((i & ~(HCJ_LUT_SUB_SIZE - 1)) | (int)sub_source_index((uint32_t)i & (HCJ_LUT_SUB_SIZE - 1), (uint32_t)l.max_bits))], k == 0);
    }
    Tables &t = ht->tab[c];
    t.dc_off = (uint32_t)(c * 2) * HCJ_LUT_SIZE;
    t.ac_off = (uint32_t)(c * 2 + 1) * HCJ_LUT_SIZE;
    t.dc_max_bits = ht->luts[c * 2].max_bits;
    t.ac_max_bits = ht->luts[c * 2 + 1].max_bits;
  }
  ht->multi.assign((size_t)plan->info.ncomp * HCJ_LUT_SIZE, 0);
  for (int c = 0; c < plan->info.ncomp; c++)  // build_multi_tables in hcj_kernels.cu
    for (uint32_t i = 0; i < HCJ_LUT_SIZE; i++)
      ht->multi[(size_t)c * HCJ_LUT_SIZE + i] = multi_sync_entry(ht->fast.data() + (size_t)(c * 2 + 1) * HCJ_LUT_SIZE, i);
  for (int k = 0; k < plan->info.blocks_per_mcu; k++) {
    const int c = plan->blk_comp[k];
    ht->blk_comp[k] = (uint8_t)c;
    uint32_t qmax = 0;
    for (int e = 1; e < 64; e++) qmax = std::max(qmax, (uint32_t)ht->quant[c * 128 + e]);
    ht->blkinfo[k] = BlkInfo{ht->tab[c].dc_off, ht->tab[c].ac_off, (uint32_t)c * 128u, (uint32_t)c | (qmax << 8)};
  }
  return 0;
}

// ---- single-lane stand-ins for the warp-synchronous fast passes of hcj_kernels.cu (warp_exact_fast,
// warp_sync_fast): the same step functions in the same order, with the warp votes removed.  A lane runs
// the fast steps until it has to leave, then hands its state to the literal loop.
struct ExactState {
  uint32_t p, cz, share;
  int64_t blk;
  int32_t pred[HCJ_MAX_COMP];
};

void store_sparse(const int16_t *row, int16_t *blk) {
  for (int i = 0; i < 64; i++)
    if (row[i]) blk[i] = row[i];
}

// Returns 0, or the status the fast pass itself reports (a run past coefficient 63), with *err_pos set.
int fast_exact_lane(const ScanCtx &sc, const FastTables T, uint32_t hi, uint32_t end_bits, int64_t nblocks_end,
                    int16_t *coefs, ExactState &st, uint32_t *err_pos) {
  const uint32_t lim = end_bits >= 32u ? end_bits - 32u : 0u;
  if (st.p >= lim) return 0;
  ExactLane s;
  s.br.init(sc.words, st.p);
  s.c = st.cz >> 8;
  s.z = st.cz & 0xffu;
  s.blk = (int32_t)st.blk;
  s.p0 = st.pred[0], s.p1 = st.pred[1], s.p2 = st.pred[2], s.p3 = st.pred[3];
  s.share = st.share;
  s.sumabs = 0;
  exact_bind_block(s, T);
  int16_t row[64];
  memset(row, 0, sizeof(row));
  for (;;) {
    if (s.z == 0u) {
      if (s.blk + 1 >= nblocks_end || s.br.pos >= lim || (s.c == 0u && s.br.pos >= hi)) break;
      if (!exact_dc_step(s, T, row)) break;
    } else {
      if (s.br.pos >= lim) break;
      exact_ac_step(s, T, row);
      if (z_block_done(s.z)) {
        if (z_no_code(s.z)) {
          exact_ac_undo_no_code(s);
          break;
        }
        if (z_overrun(s.z)) {
          *err_pos = s.br.pos;
          return HCJ_DEV_COEF_INDEX;
        }
        if (exact_share_may_be_wide(s) && exact_share(s, T, row) >= (uint32_t)HCJ_WIDE_SHARE) flag_wide_block(sc, s.blk);
        memcpy(coefs + (int64_t)s.blk * 64, row, sizeof(row));
        memset(row, 0, sizeof(row));
        exact_next_block(s, T, sc.bpm);
      }
    }
  }
  if (s.z != 0u) {  // the literal loop stores straight to memory (in the kernel the block stays staged)
    s.share = exact_share(s, T, row);
    memcpy(coefs + (int64_t)s.blk * 64, row, sizeof(row));
  }
  exact_save_pred(s);
  st.p = s.br.pos;
  st.cz = (s.c << 8) | s.z;
  st.share = s.share;
  st.blk = s.blk;
  st.pred[0] = s.p0, st.pred[1] = s.p1, st.pred[2] = s.p2, st.pred[3] = s.p3;
  return 0;
}

// subseq_sync with the fast steps in front
void fast_subseq_sync(const ScanCtx &sc, const Local L, uint32_t p, uint32_t cz, uint32_t hi, SubResult &r, uint32_t end_bits, int32_t *dpre) {
  const uint32_t lim = std::min(hi, end_bits >= 32u ? end_bits - 32u : 0u);
  const uint32_t lim_m = std::min(lim, hi >= (uint32_t)HCJ_LUT_BITS ? hi - (uint32_t)(HCJ_LUT_BITS - 1) : 0u);
  SyncLane s;
  s.br.init(sc.words, p);
  s.c = cz >> 8;
  s.z = cz & 0xffu;
  s.nstart = 0;
  s.d0 = s.d1 = s.d2 = s.d3 = 0;
  s.first_p = 0xffffffffu;
  s.nbefore = 0;
  sync_bind_block(s, L.ft);
  for (;;) {
    if (s.br.pos >= lim) break;
    if (s.z == 0u) {
      sync_dc_step(s, L.ft, dpre);
    } else {
      sync_ac_step_multi(s, L.ft, lim_m);
      if (z_block_done(s.z)) sync_next_block(s, L.ft, sc.bpm);
    }
  }
  int32_t dpre2[HCJ_MAX_COMP];
  subseq_sync(sc, L, s.br.pos, (s.c << 8) | s.z, hi, r, end_bits, dpre2);
  if (s.first_p != 0xffffffffu) {
    r.first_p = s.first_p, r.nbefore = s.nbefore;
  } else if (r.first_p != 0xffffffffu) {
    r.nbefore += s.nstart;
    dpre[0] = s.d0 + dpre2[0], dpre[1] = s.d1 + dpre2[1], dpre[2] = s.d2 + dpre2[2], dpre[3] = s.d3 + dpre2[3];
  }
  r.nstart += s.nstart;
  r.dcsum[0] += s.d0;
  r.dcsum[1] += s.d1;
  r.dcsum[2] += s.d2;
  r.dcsum[3] += s.d3;
}

// subseq_write with the fast steps in front: a run of whole MCUs from the MCU boundary at p (c = 0)
int fast_subseq_write(const ScanCtx &sc, const Local L, uint32_t p, uint32_t c, uint32_t hi, uint32_t end_bits, int64_t blk,
                      int32_t pred[HCJ_MAX_COMP], int64_t nblocks, int16_t *coefs, uint32_t *err_pos) {
  ExactState st;
  st.p = p;
  st.cz = c << 8;
  st.share = 0;
  st.blk = blk;
  for (int k = 0; k < HCJ_MAX_COMP; k++) st.pred[k] = pred[k];
  int err = fast_exact_lane(sc, L.ft, hi, end_bits, nblocks, coefs, st, err_pos);
  if (err) return err;
  return subseq_write(sc, L, st.p, st.cz >> 8, hi, end_bits, st.blk, st.pred, nblocks, coefs, err_pos, st.cz & 0xffu, st.share);
}

}  // namespace

extern "C" {

// k_idct / k_idct_blocks body.
void emu_reconstruct_blocks(const int16_t *coefs, int64_t n, const uint16_t *qt, int force_wide, uint8_t *out) {
  for (int64_t i = 0; i < n; i++) reconstruct_block(coefs + i * 64, qt, force_wide != 0, out + i * 64);
}

// Same, but forcing the 32-bit path regardless of the guard: used to probe where int32 stops being exact.
void emu_idct32_unguarded(const int32_t *dequant_natural, int64_t n, int32_t *out) {
  for (int64_t i = 0; i < n; i++) {
    int32_t v[64];
    memcpy(v, dequant_natural + i * 64, sizeof(v));
    idct_8x8<int32_t>(v);
    memcpy(out + i * 64, v, sizeof(v));
  }
}

int emu_idct_l1_limit() { return HCJ_IDCT_L1_LIMIT; }

// k_huff_restart: one "thread" per segment.  `entropy` = destuffed bytes, seg_off[nseg + 1] byte offsets.
int emu_decode_segments(const uint8_t *jpeg, int64_t len, unsigned flags, const uint8_t *entropy, const uint32_t *seg_off,
                        uint32_t nseg, int16_t *coefs /* zeroed, nblocks * 64 */, uint32_t *wide_flags /* zeroed */) {
  hcj_header *h = new hcj_header;
  hcj::ImagePlan plan;
  HostTables ht;
  int st = prepare(jpeg, (size_t)len, flags, h, &plan, &ht);
  delete h;
  if (st) return st;
  const hcj_frame_info &f = plan.info;
  uint32_t nmcu = (uint32_t)(f.mcus_wide * f.mcus_high), bpm = (uint32_t)f.blocks_per_mcu;
  uint32_t ri = f.restart_interval ? (uint32_t)f.restart_interval : nmcu;
  std::vector<uint32_t> words((seg_off[nseg] + 15) / 4 + 4, 0);
  memcpy(words.data(), entropy, seg_off[nseg]);
  const Local L = ht.local();
  ScanCtx sc;
  sc.words = words.data();
  sc.total_bits = 0;
  sc.bpm = bpm;
  for (int c = 0; c < f.ncomp; c++) sc.tab[c] = ht.tab[c];
  sc.wide_flags = wide_flags;
  sc.blk_base = 0;
  unsigned long long err_key = ~0ull;
  for (uint32_t seg = 0; seg < nseg; seg++) {  // <- thread index
    uint32_t seg_bits = (seg_off[seg + 1] - seg_off[seg]) * 8;
    uint32_t mcu0 = seg * ri, mcu1 = mcu0 + ri < nmcu ? mcu0 + ri : nmcu;
    int err = 0;
    uint32_t err_pos = 0;
    if (seg_bits > 16) {
      int32_t pred[HCJ_MAX_COMP] = {0, 0, 0, 0};
      err = fast_subseq_write(sc, L, seg_off[seg] * 8, 0, 0xffffffffu, seg_off[seg + 1] * 8, (int64_t)mcu0 * bpm - 1, pred,
                              (int64_t)mcu1 * bpm, coefs, &err_pos);
    } else {
      BitReader br;
      br.init(words.data(), seg_off[seg] * 8, seg_off[seg + 1] * 8);
      int32_t pred[HCJ_MAX_COMP] = {0, 0, 0, 0};
      int64_t blk = (int64_t)mcu0 * bpm;
      for (uint32_t mcu = mcu0; mcu < mcu1 && !err; mcu++)
        for (uint32_t k = 0; k < bpm && !err; k++, blk++) {
          int comp = plan.blk_comp[k];
          err = decode_block_exact(br, L, ht.tab[comp], seg_bits, pred[comp], coefs + blk * 64);
          flag_wide_block(sc, blk);
          err_pos = br.pos;
        }
    }
    unsigned long long key = ((unsigned long long)err_pos << 8) | (unsigned long long)(-err);
    if (err && key < err_key) err_key = key;
  }
  return err_key == ~0ull ? 0 : -(int)(err_key & 0xff);
}

// k_spec_*: one unit of the speculative decoder = a scan without restart markers, or one (long) restart interval:
// bits [lo_bits, end_bits) of the image's entropy data hold blocks [blk0, blk_end), DC predictors 0 at the start.
// T emulated threads per chunk, subsequences of S bits.
static void spec_unit(const ScanCtx &sc, const Local LT, const uint8_t *blk_comp, uint32_t lo_bits, uint32_t end_bits, int64_t blk0,
                      int64_t blk_end, int T, uint32_t S, int16_t *coefs, unsigned long long *err_key, int *max_rounds) {
  const uint32_t L = end_bits - lo_bits;
  if (L <= 16) {
    BitReader br;
    br.init(sc.words, lo_bits, end_bits);
    int32_t pred[HCJ_MAX_COMP] = {0, 0, 0, 0};
    for (int64_t blk = blk0; blk < blk_end; blk++) {
      int comp = blk_comp[blk % sc.bpm];
      int err = decode_block_exact(br, LT, sc.tab[comp], L, pred[comp], coefs + blk * 64);
      flag_wide_block(sc, blk);
      if (err) {
        unsigned long long key = ((unsigned long long)br.pos << 8) | (unsigned long long)(-err);
        if (key < *err_key) *err_key = key;
        return;
      }
    }
    return;
  }
  struct Carry {
    uint32_t p = 0, cz = 0;
    int64_t nstart = 0;
    int32_t dc[HCJ_MAX_COMP] = {0, 0, 0, 0};
  } carry;
  carry.p = lo_bits;
  carry.nstart = blk0;
  const uint32_t nsub = (L + S - 1) / S;
  for (uint32_t base = 0; base < nsub; base += T) {
    const int nact = (int)(nsub - base < (uint32_t)T ? nsub - base : (uint32_t)T);
    std::vector<SubResult> r(nact);
    std::vector<int32_t> dpre(nact * HCJ_MAX_COMP);
    std::vector<uint32_t> sp(nact), scz(nact), endp(nact), endcz(nact), hi(nact);
    // phase A
    for (int t = 0; t < nact; t++) {
      uint32_t i = base + t, lo = lo_bits + i * S;
      hi[t] = lo + S < end_bits ? lo + S : end_bits;
      sp[t] = t == 0 ? carry.p : lo;
      scz[t] = t == 0 ? carry.cz : 0;
      fast_subseq_sync(sc, LT, sp[t], scz[t], hi[t], r[t], end_bits, &dpre[t * HCJ_MAX_COMP]);
      endp[t] = r[t].p;
      endcz[t] = r[t].cz;
    }
    // phase B (Jacobi rounds: all threads read the ends of the previous round)
    int rounds = 0;
    for (;;) {
      std::vector<uint32_t> np(endp), ncz(endcz);
      bool any = false;
      int redo_cnt = 0;
      for (int t = 0; t < nact; t++) {
        uint32_t nsp = t == 0 ? carry.p : endp[t - 1], nscz = t == 0 ? carry.cz : endcz[t - 1];
        if (nsp != sp[t] || nscz != scz[t]) {
          sp[t] = nsp;
          scz[t] = nscz;
          redo_cnt++;
          fast_subseq_sync(sc, LT, nsp, nscz, hi[t], r[t], end_bits, &dpre[t * HCJ_MAX_COMP]);
          if (r[t].p != endp[t] || r[t].cz != endcz[t]) {
            any = true;
            np[t] = r[t].p;
            ncz[t] = r[t].cz;
          }
        }
      }
      endp = np;
      endcz = ncz;
      rounds++;
      if (getenv("EMU_STATS")) fprintf(stderr, "round %d redo %d of %d\n", rounds, redo_cnt, nact);
      if (!any) break;
    }
    if (rounds > *max_rounds) *max_rounds = rounds;
    // phase C: the exact pass; every thread decodes the blocks that begin in its subsequence (emulated threads run in
    // REVERSE order: nothing may depend on a left neighbour having run first)
    std::vector<int64_t> exn(nact);
    std::vector<int32_t> exd(nact * HCJ_MAX_COMP);
    {
      int64_t acc = 0;
      int32_t accd[HCJ_MAX_COMP] = {0, 0, 0, 0};
      for (int t = 0; t < nact; t++) {
        exn[t] = acc;
        for (int k = 0; k < HCJ_MAX_COMP; k++) exd[t * HCJ_MAX_COMP + k] = accd[k];
        acc += r[t].nstart;
        for (int k = 0; k < HCJ_MAX_COMP; k++) accd[k] += r[t].dcsum[k];
      }
    }
    int64_t ex_n = 0;
    int32_t ex_dc[HCJ_MAX_COMP] = {0, 0, 0, 0};
    for (int tt = 0; tt < nact; tt++) {
      const int t = nact - 1 - tt;
      ex_n = exn[t];
      for (int k = 0; k < HCJ_MAX_COMP; k++) ex_dc[k] = exd[t * HCJ_MAX_COMP + k];
      if (r[t].first_p == 0xffffffffu) continue;  // no MCU begins here
      int32_t pred[HCJ_MAX_COMP];
      for (int k = 0; k < HCJ_MAX_COMP; k++) pred[k] = carry.dc[k] + ex_dc[k] + dpre[t * HCJ_MAX_COMP + k];
      int64_t blk = carry.nstart + ex_n + r[t].nbefore - 1;
      uint32_t err_pos = 0;
      int err = fast_subseq_write(sc, LT, r[t].first_p, 0, hi[t], end_bits, blk, pred, blk_end, coefs, &err_pos);
      unsigned long long key = ((unsigned long long)err_pos << 8) | (unsigned long long)(-err);
      if (err && key < *err_key) *err_key = key;  // the kernel's atomicMin
    }
    ex_n = exn[nact - 1] + r[nact - 1].nstart;
    for (int k = 0; k < HCJ_MAX_COMP; k++) ex_dc[k] = exd[(nact - 1) * HCJ_MAX_COMP + k] + r[nact - 1].dcsum[k];
    carry.p = r[nact - 1].p;
    carry.cz = r[nact - 1].cz;
    carry.nstart += ex_n;
    for (int k = 0; k < HCJ_MAX_COMP; k++) carry.dc[k] += ex_dc[k];
  }
}

// k_spec_*: a scan without restart markers (nseg = 1, seg_off = {0, ent_len}) or one whose restart intervals are long
// enough to be decoded speculatively, each interval a unit of its own.  Returns status; rounds_out gets the largest
// number of fix-point rounds any chunk needed.
int emu_decode_units(const uint8_t *jpeg, int64_t len, unsigned flags, const uint8_t *entropy, const uint32_t *seg_off, uint32_t nseg,
                     int T, uint32_t S, int16_t *coefs /* zeroed */, int *rounds_out, uint32_t *wide_flags /* zeroed, nblocks / 32 + 1 */) {
  hcj_header *h = new hcj_header;
  hcj::ImagePlan plan;
  HostTables ht;
  int st = prepare(jpeg, (size_t)len, flags, h, &plan, &ht);
  delete h;
  if (st) return st;
  const hcj_frame_info &f = plan.info;
  const uint32_t ent_len = seg_off[nseg];
  std::vector<uint32_t> words((ent_len + 15) / 4 + 4, 0);
  memcpy(words.data(), entropy, ent_len);
  const Local LT = ht.local();
  ScanCtx sc;
  sc.words = words.data();
  sc.total_bits = ent_len * 8;
  sc.bpm = (uint32_t)f.blocks_per_mcu;
  for (int c = 0; c < f.ncomp; c++) sc.tab[c] = ht.tab[c];
  sc.wide_flags = wide_flags;
  sc.blk_base = 0;
  const uint32_t nmcu = (uint32_t)(f.mcus_wide * f.mcus_high);
  const uint32_t ri = nseg > 1 || f.restart_interval ? (uint32_t)f.restart_interval : nmcu;
  int max_rounds = 0;
  unsigned long long err_key = ~0ull;
  for (uint32_t u = 0; u < nseg; u++) {
    const uint32_t mcu0 = std::min(u * ri, nmcu), mcu1 = std::min(mcu0 + ri, nmcu);
    spec_unit(sc, LT, ht.blk_comp, seg_off[u] * 8, seg_off[u + 1] * 8, (int64_t)mcu0 * sc.bpm, (int64_t)mcu1 * sc.bpm, T, S, coefs, &err_key,
              &max_rounds);
  }
  if (rounds_out) *rounds_out = max_rounds;
  return err_key == ~0ull ? 0 : -(int)(err_key & 0xff);
}

int emu_decode_speculative(const uint8_t *jpeg, int64_t len, const uint8_t *entropy, uint32_t ent_len, int T, uint32_t S,
                           int16_t *coefs /* zeroed */, int *rounds_out, uint32_t *wide_flags /* zeroed, nblocks / 32 + 1 */) {
  const uint32_t seg_off[2] = {0, ent_len};
  return emu_decode_units(jpeg, len, 0, entropy, seg_off, 1, T, S, coefs, rounds_out, wide_flags);
}

// k_fdct_quant body for one 8x8 block of pixels (row-major), quant table in zig-zag order.
void emu_fdct_quant(const uint8_t *pix, const uint16_t *qt, int16_t *out_zigzag, int32_t *out_fdct) {
  int32_t v[64];
  for (int i = 0; i < 64; i++) v[i] = (int32_t)pix[i] - 128;
  fdct_8x8(v);
  if (out_fdct) memcpy(out_fdct, v, sizeof(v));
  for (int z = 0; z < 64; z++) {
    uint32_t recip = (uint32_t)((1ull << 32) / (4u * qt[z])) + 1u;
    out_zigzag[z] = (int16_t)quantize(v[zigzag_inverse(z)], qt[z], recip);
  }
}

// Exhaustive check of the reciprocal quantiser against truncating division; returns mismatches.
int64_t emu_quantize_check(int fmax) {
  int64_t bad = 0;
  for (uint32_t q = 1; q <= 255; q++) {
    uint32_t recip = (uint32_t)((1ull << 32) / (4u * q)) + 1u;
    for (int f = -fmax; f <= fmax; f++) {
      int32_t want = f < 0 ? (f - (int)q * 2) / ((int)q * 4) : (f + (int)q * 2) / ((int)q * 4);
      if (quantize(f, q, recip) != want) bad++;
    }
  }
  return bad;
}

// k_block_bits / k_pack symbol stream for a whole frame of quantised blocks -> entropy-coded bytes with
// stuffing, 1-fill and (optionally) RSTn, i.e. everything between the SOS header and EOI.
struct FieldLog {  // the bits a block emits, one per element (fields may be merged or split differently)
  std::vector<uint8_t> f;
  void operator()(uint32_t bits, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) f.push_back((uint8_t)((bits >> (n - 1 - i)) & 1u));
  }
};
struct ByteWriter {
  std::vector<uint8_t> *out;
  uint64_t acc = 0;
  int nacc = 0;
  void operator()(uint32_t bits, uint32_t n) {
    if (!n) return;
    acc = (acc << n) | (bits & ((1u << n) - 1u));
    nacc += (int)n;
    while (nacc >= 8) {
      uint8_t b = (uint8_t)(acc >> (nacc - 8));
      out->push_back(b);
      if (b == 0xff) out->push_back(0);
      nacc -= 8;
    }
  }
  void pad() {
    if (nacc) (*this)((1u << (8 - nacc)) - 1u, (uint32_t)(8 - nacc));
  }
};

int64_t emu_entropy_encode(const int16_t *quant, int width, int height, int chroma, int restart_interval, uint8_t *out,
                           int64_t cap) {
  hcj::EncodePlan p;
  if (hcj::plan_encode(width, height, chroma, 75, restart_interval, &p)) return -1;
  uint32_t dc[2][16], ac[2][256];
  hcj::encoder_tables(0, 2, dc[0], ac[0]);
  hcj::encoder_tables(1, 3, dc[1], ac[1]);
  std::vector<uint8_t> bytes;
  ByteWriter w;
  w.out = &bytes;
  int32_t pred[3] = {0, 0, 0};
  int64_t nmcu = (int64_t)p.mcus_wide * p.mcus_high;
  int rst = 0;
  for (int64_t mcu = 0; mcu < nmcu; mcu++) {
    if (restart_interval && mcu && mcu % restart_interval == 0) {
      w.pad();
      bytes.push_back(0xff);
      bytes.push_back((uint8_t)(0xd0 + (rst++ & 7)));
      pred[0] = pred[1] = pred[2] = 0;
    }
    for (int k = 0; k < p.bpm; k++) {
      const int16_t *q = quant + (mcu * p.bpm + k) * 64;
      int c = p.blk_comp[k];
      // the literal 63-step loop and the non-zero-map loop the kernels run must write the same fields
      uint32_t qw[32];
      for (int j = 0; j < 32; j++) qw[j] = (uint32_t)(uint16_t)q[2 * j] | ((uint32_t)(uint16_t)q[2 * j + 1] << 16);
      FieldLog a, b2;
      bool ok1 = encode_block_fields(q, (int32_t)q[0] - pred[c], dc[c ? 1 : 0], ac[c ? 1 : 0], a);
      bool ok2 = encode_block_fields_sparse(nonzero_map(qw), [q](int k) { return (int32_t)q[k]; }, (int32_t)q[0] - pred[c],
                                            dc[c ? 1 : 0], ac[c ? 1 : 0], b2);
      if (ok1 != ok2 || a.f != b2.f) return -4;
      if (!encode_block_fields_sparse(nonzero_map(qw), [q](int k) { return (int32_t)q[k]; }, (int32_t)q[0] - pred[c],
                                      dc[c ? 1 : 0], ac[c ? 1 : 0], w))
        return -2;
      pred[c] = q[0];
    }
  }
  w.pad();
  if ((int64_t)bytes.size() > cap) return -3;
  memcpy(out, bytes.data(), bytes.size());
  return (int64_t)bytes.size();
}

}  // extern "C"
