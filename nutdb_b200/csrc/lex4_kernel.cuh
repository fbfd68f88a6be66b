// lex4_kernel.cuh -- k_lex4: the single-pass lexer over RANGES of whole statements (logic: lex3_core.cuh).
//
// Statements are independent units (Parser::parse has no cross-statement state, reference src/parser/mod.rs:21-37)
// and the caller hands us their offsets.  So the batch is cut, at statement starts, into ranges of ~100 KB; a range
// begins in the code state with no literal open, and one CTA lexes it tile by tile (8 KB) with the three carries
// (context state, token count, statement start / open literal) in shared memory.  No CTA ever waits for another:
// the chained look-back scans of k_lex3 (lex3_kernels.cuh) -- which measured as a convoy: every tile waits for the
// slowest tile before it, 10-13 K of 60 K cycles per tile, whatever the look-back width -- are only needed when a
// statement is too long to cut around, and k_lex3 stays as that fallback.
//
// The price: a range's tokens go to a SEGMENT of the token arrays sized by an estimate (bytes / 4 + 2 per statement),
// so the arrays have gaps between segments.  The parser never notices (a statement's tokens are contiguous, and its
// range is found from per-window token indices as before); callers that want the token arrays get them compacted by
// k_tok_compact.
#pragma once
#include "lex3_kernels.cuh"

#ifndef L4_THREADS
#define L4_THREADS 256
#endif
static_assert(L4_THREADS == L3_WIN, "k_lex4: one thread per window of the tile");
#ifndef L4_MINBLOCKS
#define L4_MINBLOCKS 4  // (64 registers with a few spilled words: the fourth resident CTA is worth more -- measured 1.77 -> 1.69 ms per 256 MiB)
#endif

struct Lex4Ranges {
  const uint32_t* byte_begin;  // [nranges + 1] first byte of each range, relative to the batch; [nranges] = n
  const uint32_t* tok_base;    // [nranges + 1] first token slot of each range's segment; the difference is its capacity
  uint2* tok_count;            // [nranges]     .x = tokens the range produced (more than its capacity: the batch is re-run)
  const uint32_t* nranges;     // (computed on the device by k_cuts)
};

struct alignas(16) Lex4Shared {
  alignas(16) uint8_t pad[16];
  alignas(16) uint8_t text[2][L3_HALO + L3_TILE + L3_HALO];
  alignas(16) uint32_t bm[2][L3_WIN + 4];
  alignas(8) unsigned long long bar[2];
  LexTables T;
  uint32_t bndm[L3_WIN + 1];
  uint32_t sst_in[L3_WIN];
  uint32_t rec0[L3_RCAP], rec1[L3_RCAP];
  uint16_t wlist[L3_RCAP];  // records of the words that may be keywords (pass 2 of the token stage)
  uint32_t nwords;
  uint32_t wfn[L3_WARPS];
  uint4 wsum[L3_WARPS];
  uint32_t range;      // ticket
  uint32_t car_state;  // context state at the first byte of the tile
  uint4 car;           // tokens of the range so far / 1 + last statement start / open quote / flags
};

// stages [tile_begin - 32, tile_begin + 8 KB + 32) and the tile's bitmap words; tile_begin is a multiple of 128
__device__ __forceinline__ void l4_issue_load(Lex4Shared& S, int b, uint32_t tile_begin, const uint8_t* text,
                                              const uint32_t* bitmap, uint32_t n_readable) {
  const uint32_t lo = tile_begin >= L3_HALO ? tile_begin - L3_HALO : 0u;
  uint32_t hi = tile_begin + L3_TILE + L3_HALO;
  if (hi > n_readable) hi = n_readable & ~15u;  // whole 16-byte pieces only; the ragged rest is loaded by threads
  const uint32_t bytes_text = hi > lo ? hi - lo : 0u;
  const uint32_t bytes_bm = (L3_WIN + 4) * 4u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(&S.bar[b], bytes_text + bytes_bm);
  if (bytes_text) bulk_g2s(&S.text[b][L3_HALO + lo - tile_begin], text + lo, bytes_text, &S.bar[b]);
  bulk_g2s(&S.bm[b][0], bitmap + (tile_begin >> 5), bytes_bm, &S.bar[b]);
}

__global__ void __launch_bounds__(L4_THREADS, L4_MINBLOCKS) k_lex4(const uint8_t* __restrict__ text,
                                                                    const uint32_t* __restrict__ bitmap, uint32_t n,
                                                                    uint32_t n_readable, const LexTables* __restrict__ gT,
                                                                    Lex4Ranges rg, Lex3Out out, uint32_t* gate) {
  if (*gate & 1u) return;  // invalid statement offsets (k_prep): nothing downstream may trust them
  extern __shared__ __align__(16) unsigned char l3_smem[];
  Lex4Shared& S = *reinterpret_cast<Lex4Shared*>(l3_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t full = 0xFFFFFFFFu;
  stage_tables(gT, &S.T);
  if (threadIdx.x == 0) {
    mbar_init(&S.bar[0], 1);
    mbar_init(&S.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase0 = 0u, phase1 = 0u;
  int b = 0;
  for (;;) {
    if (threadIdx.x == 0) S.range = atomicAdd(out.counters + 3, 1u);
    __syncthreads();
    const uint32_t r = S.range;
    if (r >= *rg.nranges) break;
    const uint32_t rb = rg.byte_begin[r], re = rg.byte_begin[r + 1];
    const uint32_t seg_base = rg.tok_base[r], seg_cap = rg.tok_base[r + 1] - seg_base;
    const uint32_t t_first = rb & ~127u;
    const uint32_t ntile_r = re > t_first ? (re - t_first + L3_TILE - 1u) / L3_TILE : 0u;
    if (threadIdx.x == 0) {
      S.car_state = A_C;                            // a range begins with a statement: code, nothing open
      S.car = make_uint4(0u, rb + 1u, 0u, 0u);
      if (ntile_r) l4_issue_load(S, b, t_first, text, bitmap, n_readable);
    }
    __syncthreads();
    for (uint32_t k = 0; k < ntile_r; k++, b ^= 1) {
      const uint32_t tile_begin = t_first + k * L3_TILE;
      if (threadIdx.x == 0 && k + 1 < ntile_r) l4_issue_load(S, b ^ 1, tile_begin + L3_TILE, text, bitmap, n_readable);
      mbar_wait(&S.bar[b], b ? phase1 : phase0);
      if (b) phase1 ^= 1u;
      else phase0 ^= 1u;
      uint8_t* const sm = &S.text[b][L3_HALO];
      const uint32_t* const bm = S.bm[b];
      Tile3Src src{text, sm, tile_begin, n};
      // what the bulk copy could not bring: the front halo at the batch start, and everything behind the last whole
      // 16-byte piece of a caller's unpadded buffer (loaded bytewise up to n, zero beyond)
      if (tile_begin == 0 && threadIdx.x < L3_HALO) sm[(int)threadIdx.x - L3_HALO] = 0;
      if (tile_begin + L3_TILE + L3_HALO > n_readable) {
        const uint32_t from = n_readable & ~15u;
        for (uint32_t p = max(from, tile_begin >= L3_HALO ? tile_begin - L3_HALO : 0u) + threadIdx.x;
             p < tile_begin + L3_TILE + L3_HALO; p += L4_THREADS)
          sm[(int)(p - tile_begin)] = p < n ? text[p] : (uint8_t)0;
        __syncthreads();
      } else if (tile_begin == 0) {
        __syncthreads();
      }
      const uint32_t base = tile_begin + 32u * threadIdx.x;
      const uint32_t blk = tile_begin + 1024u * (uint32_t)warp;
      // bytes of this window that belong to the range
      uint32_t valid = 0;
      if (base < re && base + 32u > rb) {
        valid = full;
        if (base < rb) valid &= ~((1u << (rb - base)) - 1u);
        if (re - base < 32u) valid &= (1u << (re - base)) - 1u;
      }
      const bool live = valid != 0u;
      // ---------------- stage 1: class masks, escapes, context events, transition function ----------------
      nlex2::Win w;
      nlex3::Ops op;
      {
        const uint4* wp = reinterpret_cast<const uint4*>(sm + 32u * threadIdx.x);
        const uint4 q0 = wp[0], q1 = wp[1];
        const uint32_t v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        uint32_t p[8];
        nlex3::bit_planes(v, p);
        nlex3::classify_planes(p, valid, w, op);
      }
      {
        uint32_t bnd = bm[threadIdx.x] & valid;
        if (re >= base && re - base < 32u && base + 32u > rb) bnd |= 1u << (re - base);  // the range end terminates the last statement
        w.bnd = bnd;
      }
      nlex2::Next nx;
      if (base + 32u >= re) {
        nx.byte = 0;
        nx.bnd = 1;
        nx.cls = 0;
      } else {
        nx.byte = sm[32u * threadIdx.x + 32u];
        nx.bnd = (uint8_t)(bm[threadIdx.x + 1] & 1u);
        const uint8_t pr = S.T.prop[nx.byte];
        nx.cls = (uint16_t)(((pr & PR_IDENT_END) ? 1u : 0u) | ((pr & PR_NUM_END) ? 2u : 0u));
      }
      uint8_t prev_byte = 0, prev2_byte = 0, esc_in = 0;
      {
        const uint32_t bs_prev = __shfl_up_sync(full, w.bs, 1);
        if (live && base > rb) {
          prev_byte = sm[(int)(32u * threadIdx.x) - 1];
          if (base > rb + 1u) prev2_byte = sm[(int)(32u * threadIdx.x) - 2];
          if (lane == 0) {
            if (!(w.bnd & 1u) && prev_byte == '\\') {  // backslash parity in front of the warp's block: walk back over the run
              uint32_t nrun = 0, p = base;
              while (p > rb && src.byte(p - 1) == '\\') {
                nrun++;
                p--;
                if ((bitmap[p >> 5] >> (p & 31u)) & 1u) break;
              }
              esc_in = (uint8_t)(nrun & 1u);
            }
          } else {
            const int run = nlex2::clz32(~bs_prev);
            esc_in = (uint8_t)(run >= 32 ? 0 : (run & 1));  // (32 backslashes in a row: the statement is flagged below)
          }
        }
      }
      const uint32_t escm = nlex2::esc_mask32(w.bs, esc_in) & ~w.bnd;
      const nlex2::Events ev = nlex2::make_events(w, escm, prev_byte);
      uint32_t fnv = NUTDB_VEC8_ID;
      if (live)
        fnv = ev.all ? nlex2::ctx_window_fn(S.T, w, ev, NUTDB_VEC8_ID)
                     : vec8_then_row(NUTDB_VEC8_ID, S.T.a_row[EV_OTHER][0], S.T.a_row[EV_OTHER][1]);
      uint32_t fexcl;
      {
        const uint32_t incl = warp_scan_vec8(fnv, lane, fexcl);
        if (lane == 31) S.wfn[warp] = incl;
      }
      if (threadIdx.x == 0)  // the byte behind the tile starts a statement (or is the end of the range)
        S.bndm[L3_WIN] = (tile_begin + L3_TILE < re ? (bm[L3_WIN] & 1u) : 0u) | (tile_begin + L3_TILE == re ? 1u : 0u);
      S.bndm[threadIdx.x] = w.bnd;
      __syncthreads();  // ---- A: wfn
      uint8_t s_warp;
      {
        uint32_t st = S.car_state;  // (applying the warps' functions one after the other is far cheaper than composing them)
        for (int i = 0; i < warp; i++) st = vec8_apply(S.wfn[i], st);
        s_warp = (uint8_t)st;
      }
      const uint8_t s_in = (uint8_t)vec8_apply(fexcl, s_warp);
      // ---------------- stage 2: concrete context walk, token masks ----------------
      nlex3::WinCtx3 o;
      if (live) nlex3::ctx_window3(w, ev, base, nx, s_in, prev_byte, o);
      o.escm = escm;
      nlex3::Hist3 h;
      {
        h.L = __shfl_up_sync(full, w.L & o.ct, 1);
        h.D = __shfl_up_sync(full, w.D & o.ct, 1);
        h.DOT = __shfl_up_sync(full, w.DOT & o.ct, 1);
        h.bnd = __shfl_up_sync(full, w.bnd, 1);
        // lane 0: the 32 bytes in front of the warp's block; their raw classes are exact where it matters (a run of
        // word characters / dots that reaches the block) when the block is entered in code
        uint32_t rL = 0, rD = 0, rDOT = 0;
        if (blk >= 32u) {
          const uint8_t c = sm[(int)(blk - tile_begin) - 32 + lane];
          const uint8_t pr = S.T.prop[c];
          rL = __ballot_sync(full, (pr & PR_WORD) && !(pr & PR_DIGIT));
          rD = __ballot_sync(full, pr & PR_DIGIT);
          rDOT = __ballot_sync(full, c == '.');
        }
        if (lane == 0) {
          const bool code_entry = s_warp <= A_CX && blk > rb;
          h.L = code_entry ? rL : 0u;
          h.D = code_entry ? rD : 0u;
          h.DOT = code_entry ? rDOT : 0u;
          h.bnd = blk >= 32u ? (warp == 0 ? bitmap[(blk - 32u) >> 5] : bm[(blk - 32u - tile_begin) >> 5]) : 0u;
        }
      }
      nlex3::TokMasks m;
      uint4 mine = make_uint4(0u, 0u, 0u, 0u);
      uint32_t bad_carry = 0;
      if (live) {
        nlex3::win_tokens3(w, op, o, h, nx, prev_byte, prev2_byte, m);
        uint32_t bad = m.bad | nlex2::win_bad_mask(src, w, o, base, prev_byte);
        if (w.bs == 0xFFFFFFFFu) bad |= 1u;  // a backslash run longer than a window: parity not tracked
        // flag statements: the start of the statement each flagged byte belongs to (the one open at the window
        // start is known once the carries are scanned)
        uint32_t bb = bad;
        const uint32_t bnds = w.bnd & valid;
        while (bb) {
          const int i = __ffs((int)bb) - 1;
          bb &= bb - 1;
          const uint32_t below = bnds & (i >= 31 ? 0xFFFFFFFFu : ((2u << i) - 1u));
          if (below) out.punt_stmt_at(base + (uint32_t)(31 - __clz((int)below)));
          else bad_carry = 1;
        }
        bb = o.bad_prev;  // the statement that ENDS right before this (statement start) byte
        while (bb) {
          const int i = __ffs((int)bb) - 1;
          bb &= bb - 1;
          const uint32_t below = bnds & ((1u << i) - 1u);
          if (below) out.punt_stmt_at(base + (uint32_t)(31 - __clz((int)below)));
          else bad_carry = 1;
        }
        // the range ends exactly on a window boundary inside a literal / comment: no window carries the virtual start
        if (base + 32u == re && (o.s_out == A_SQ || o.s_out == A_DQ || o.s_out == A_BT || o.s_out == A_BC0 || o.s_out == A_BC)) {
          if (bnds) out.punt_stmt_at(base + (uint32_t)(31 - __clz((int)bnds)));
          else bad_carry = 1;
        }
        mine.x = (uint32_t)__popc(m.has) + (uint32_t)__popc(m.eofm);
        mine.y = o.last_bnd1;
        mine.z = o.sc.open_pos;
        mine.w = (uint32_t)(o.sc.has_open != 0) | ((uint32_t)(o.sc.esc != 0) << 1) | ((uint32_t)(o.sc.chk != 0) << 3);
      }
      // ---------------- scan of (count, statement start, open literal) over the tile ----------------
      uint4 incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint4 o2 = c3_shfl_up(incl, d);
        if (lane >= d) incl = c3_then(o2, incl);
      }
      uint4 excl = c3_shfl_up(incl, 1);
      if (lane == 0) excl = make_uint4(0u, 0u, 0u, 0u);
      if (lane == 31) S.wsum[warp] = incl;
      __syncthreads();  // ---- C: wsum
      uint4 cin = S.car;  // everything of the range before this thread's window
      uint32_t tile_count = 0;
      {
        for (int i = 0; i < L3_WARPS; i++) {
          if (i == warp) cin = c3_then(cin, excl);
          else if (i < warp) cin = c3_then(cin, S.wsum[i]);
          tile_count += S.wsum[i].x;
        }
      }
      const uint32_t tile_first = S.car.x;  // tokens of the range before the tile
      // the carries behind this tile (thread 0 writes them once everybody is done with the old ones)
      uint32_t new_state = 0;
      uint4 new_car = make_uint4(0u, 0u, 0u, 0u);
      if (threadIdx.x == 0) {
        new_car = S.car;
        new_state = S.car_state;
#pragma unroll
        for (int i = 0; i < L3_WARPS; i++) {
          new_state = vec8_apply(S.wfn[i], new_state);
          new_car = c3_then(new_car, S.wsum[i]);
        }
      }
      const uint32_t sst_open = cin.y ? cin.y - 1u : 0u;
      S.sst_in[threadIdx.x] = sst_open;
      if (live && bad_carry) out.punt_stmt_at(sst_open);
      if (base <= re && base + 32u > rb) {  // (the window that starts exactly at the range end holds the last statement's token-range end)
        const size_t slot = (size_t)(base >> 5) + r;  // one extra slot per range: two ranges may share a window
        out.win_idx[slot] = seg_base + cin.x;
        out.win_has[slot] = m.has;
        out.win_eof[slot] = m.eofm;
      }
      // ---------------- stage 3: token records (thread per window) -> tokens (thread per token) ----------------
      nlex2::StrCarry sc_in;
      sc_in.has_open = (uint8_t)(cin.w & 1u);
      sc_in.esc = (uint8_t)((cin.w >> 1) & 1u);
      sc_in.chk = (uint8_t)((cin.w >> 3) & 1u);
      sc_in.open_pos = cin.z;
      const uint32_t local = cin.x - tile_first;
      for (uint32_t r0 = 0; r0 < tile_count; r0 += L3_RCAP) {
        if (r0) __syncthreads();  // the previous round's records have been consumed
        if (live && mine.x && local < r0 + L3_RCAP && local + mine.x > r0) {
          auto rec = [&](uint32_t idx, uint32_t start_abs, uint32_t end_abs, uint32_t flags) {
            const uint32_t q = idx - r0;
            if (q < L3_RCAP) {
              S.rec0[q] = start_abs;
              S.rec1[q] = (end_abs - tile_begin) | (flags << NUTDB_R3_KIND_SHIFT);
            }
          };
          nlex3::win_records3(o, m, base, sc_in, local, rec);
        }
        __syncthreads();  // ---- E: records
        const uint32_t cnt = min(tile_count - r0, (uint32_t)L3_RCAP);
        if (threadIdx.x == 0) S.nwords = 0u;
        __syncthreads();
        // pass 1: every token -- type, statement-relative span; words that can be keywords are listed for pass 2
        for (uint32_t q0 = 0; q0 < cnt; q0 += L4_THREADS) {
          const uint32_t q = q0 + threadIdx.x;
          bool is_word = false;
          if (q < cnt) {
            const uint32_t start_abs = S.rec0[q], r1 = S.rec1[q];
            const uint32_t end_rel = r1 & ((1u << NUTDB_R3_KIND_SHIFT) - 1u), flags = r1 >> NUTDB_R3_KIND_SHIFT;
            const uint32_t last = end_rel - 1u;  // the token's last byte, tile relative
            const uint32_t wv = last >> 5, i = last & 31u;
            const uint32_t bb = S.bndm[wv] & (i >= 31u ? 0xFFFFFFFFu : ((2u << i) - 1u));
            const uint32_t sst = bb ? tile_begin + 32u * wv + (uint32_t)(31 - __clz((int)bb)) : S.sst_in[wv];
            nlex3::Tok3 tk;
            is_word = nlex3::token_finish3(S.T, src, start_abs, tile_begin + end_rel, flags, sst, tk);
            if (tk.punt) {
              out.punt_stmt_at(sst);
              tk.type = NUTDB_TT_POISON;  // (the slot belongs to a flagged statement: a fixed filler)
              tk.start = tk.end = 0;
              tk.kw = 0;
            }
            const uint32_t li = tile_first + r0 + q;  // index inside the range's segment
            if (li < seg_cap) {
              const uint32_t gi = seg_base + li;
              out.type[gi] = tk.type;
              out.start[gi] = tk.start;
              out.end[gi] = tk.end;
              if (!is_word) out.kw[gi] = tk.kw;
            } else {
              is_word = false;
            }
          }
          const uint32_t wm = __ballot_sync(full, is_word);
          if (wm) {
            uint32_t wbase = 0;
            if (lane == 0) wbase = atomicAdd(&S.nwords, (uint32_t)__popc(wm));
            wbase = __shfl_sync(full, wbase, 0);
            if (is_word) S.wlist[wbase + (uint32_t)__popc(wm & ((1u << lane) - 1u))] = (uint16_t)q;
          }
        }
        __syncthreads();
        // pass 2: keyword ids (perfect hash), all lanes on the same code
        const uint32_t nw = S.nwords;
        for (uint32_t j = threadIdx.x; j < nw; j += L4_THREADS) {
          const uint32_t q = S.wlist[j];
          const uint32_t start_abs = S.rec0[q], end_abs = tile_begin + (S.rec1[q] & ((1u << NUTDB_R3_KIND_SHIFT) - 1u));
          const uint32_t len = end_abs - start_abs;
          uint32_t w0, w1, w2;
          load_word12(src.span(start_abs, len), len, w0, w1, w2);
          out.kw[seg_base + tile_first + r0 + q] = keyword_lookup_words(S.T, len, w0, w1, w2);
        }
      }
      __syncthreads();  // ---- F: everyone is done with buffer b, the per-window tables and the carries
      if (threadIdx.x == 0) {
        S.car_state = new_state;
        S.car = new_car;
        if (k + 1 == ntile_r) {
          rg.tok_count[r] = make_uint2(new_car.x, 0u);
          // (a range that ends on a window boundary: that window -- the last statement's token-range end -- may lie
          // behind the last tile)
          if ((re & 31u) == 0u) out.win_idx[(size_t)(re >> 5) + r] = seg_base + new_car.x;
          if (new_car.x > seg_cap) {  // the segment was too small: the batch is run again, and the kernels behind
            atomicOr(out.counters + 4, 1u);  // this one skip the attempt (bit 1 of the gate): its token ranges
            atomicOr(gate, 2u);              // run into the neighbouring segments
          }
        }
      }
      // (the carries are read after barrier A / C of the next tile, or after the barrier that follows the next ticket)
    }
  }
}

// ---- the ranges: cut points at statement starts nearest to the multiples of `target` bytes --------------------
// One block.  cut r = the first non-empty statement that starts at or behind byte r * target; duplicates (a statement
// longer than `target`) are dropped.  Outputs: byte_begin / cut_stmt / tok_base [nranges + 1], info = {nranges,
// end of the last segment, longest range}.  Token capacity of a range = bytes * cap_num / 16 + 2 per statement + 64.
#define L4_MAX_RANGES 2048
__global__ void __launch_bounds__(1024) k_cuts(const uint32_t* __restrict__ off32, uint32_t nstmt, uint32_t n, uint32_t target,
                                               uint32_t cap_num, uint32_t* __restrict__ byte_begin, uint32_t* __restrict__ cut_stmt,
                                               uint32_t* __restrict__ tok_base, uint32_t* __restrict__ info,
                                               const uint32_t* __restrict__ gate) {
  __shared__ uint32_t s_stmt[L4_MAX_RANGES + 1], s_keep[L4_MAX_RANGES + 1], s_pos[L4_MAX_RANGES + 2];
  __shared__ uint32_t s_n;
  if (*gate) {
    if (threadIdx.x == 0) info[0] = 0u;
    return;
  }
  const uint32_t want = min((uint32_t)L4_MAX_RANGES, (n + target - 1u) / target);
  for (uint32_t r = threadIdx.x; r < want; r += 1024u) {
    const uint32_t t = r * target;
    uint32_t lo = 0, hi = nstmt;  // first statement with off32 >= t
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (off32[mid] < t) lo = mid + 1u;
      else hi = mid;
    }
    while (lo < nstmt && off32[lo + 1] == off32[lo]) lo++;  // (empty statements own no byte)
    s_stmt[r] = lo;
  }
  __syncthreads();
  for (uint32_t r = threadIdx.x; r < want; r += 1024u)
    s_keep[r] = (s_stmt[r] < nstmt && (r == 0 || s_stmt[r] != s_stmt[r - 1])) ? 1u : 0u;
  __syncthreads();
  __shared__ uint32_t s_cap[L4_MAX_RANGES + 1], s_b0[L4_MAX_RANGES + 1], s_long;
  __shared__ uint32_t s_wtot[32];
  // exclusive prefix sum of v over the block's 2 x 1024 slots (slot = 2 * thread + {0, 1}); returns this thread's
  // exclusive prefix for its first slot and the grand total
  auto block_excl2 = [&](uint32_t v0, uint32_t v1, uint32_t& total) -> uint32_t {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = v0 + v1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= (uint32_t)d) incl += o;
    }
    if (lane == 31u) s_wtot[warp] = incl;
    __syncthreads();
    uint32_t before = 0, tot = 0;
    for (uint32_t q = 0; q < 32u; q++) {
      const uint32_t w = s_wtot[q];
      if (q < warp) before += w;
      tot += w;
    }
    __syncthreads();
    total = tot;
    return before + incl - (v0 + v1);
  };
  static_assert(L4_MAX_RANGES == 2048, "k_cuts: two slots per thread of its 1024-thread block");
  {  // compaction of the kept cut points (was a serial pass of thread 0: most of this kernel's 0.13 ms)
    const uint32_t r0 = 2u * threadIdx.x, r1 = r0 + 1u;
    const uint32_t k0 = r0 < want ? s_keep[r0] : 0u, k1 = r1 < want ? s_keep[r1] : 0u;
    uint32_t total;
    const uint32_t pos = block_excl2(k0, k1, total);
    if (k0) s_pos[pos] = s_stmt[r0];
    if (k1) s_pos[pos + k0] = s_stmt[r1];
    if (threadIdx.x == 0) {
      s_pos[total] = nstmt;
      s_n = total;
      s_long = 0;
    }
  }
  __syncthreads();
  const uint32_t k = s_n;
  for (uint32_t r = threadIdx.x; r < k; r += 1024u) {  // (the global loads, in parallel)
    const uint32_t b0 = off32[s_pos[r]], b1 = r + 1 < k ? off32[s_pos[r + 1]] : n;
    const unsigned long long cap = (unsigned long long)(b1 - b0) * cap_num / 16ull + 2ull * (s_pos[r + 1] - s_pos[r]) + 64ull;
    s_b0[r] = b0;
    s_cap[r] = (uint32_t)min(cap, (unsigned long long)(b1 - b0) + (s_pos[r + 1] - s_pos[r]) + 2ull);  // (never more than a token per byte + EOFs)
    atomicMax(&s_long, b1 - b0);
  }
  __syncthreads();
  {  // capacities -> first token slot of every range
    const uint32_t r0 = 2u * threadIdx.x, r1 = r0 + 1u;
    const uint32_t c0 = r0 < k ? s_cap[r0] : 0u, c1 = r1 < k ? s_cap[r1] : 0u;
    uint32_t total;
    const uint32_t pos = block_excl2(c0, c1, total);
    if (r0 < k) s_cap[r0] = pos;
    if (r1 < k) s_cap[r1] = pos + c0;
    if (threadIdx.x == 0) {
      s_cap[k] = total;
      s_b0[k] = n;
      info[0] = k;
      info[1] = total;
      info[2] = s_long;
    }
  }
  __syncthreads();
  for (uint32_t r = threadIdx.x; r <= k; r += 1024u) {
    byte_begin[r] = s_b0[r];
    cut_stmt[r] = s_pos[r];
    tok_base[r] = s_cap[r];
  }
}

// ---- dense token arrays for callers that want them: segment r of the lexer -> [dense_base[r], + tok_count[r]) ----
// dense_base = exclusive prefix of tok_count (k_scan_tiles), the exact lexer's region follows the last segment.
__global__ void __launch_bounds__(256) k_tok_compact(const uint8_t* __restrict__ s_type, const uint32_t* __restrict__ s_start,
                                                     const uint32_t* __restrict__ s_end, const uint8_t* __restrict__ s_kw,
                                                     uint8_t* __restrict__ d_type, uint32_t* __restrict__ d_start,
                                                     uint32_t* __restrict__ d_end, uint8_t* __restrict__ d_kw,
                                                     const uint32_t* __restrict__ tok_base, const uint2* __restrict__ tok_count,
                                                     const uint2* __restrict__ dense_base, const uint32_t* __restrict__ nranges_dev,
                                                     const uint32_t* __restrict__ n_extra_dev, uint32_t slices) {
  const uint32_t nranges = *nranges_dev;
  const uint32_t r = blockIdx.x / slices, sl = blockIdx.x % slices;  // range nranges = the exact lexer's region
  uint32_t src0, dst0, cnt;
  if (r > nranges) return;
  if (r < nranges) {
    src0 = tok_base[r];
    dst0 = dense_base[r].x;
    cnt = tok_count[r].x;
  } else {
    src0 = tok_base[nranges];
    dst0 = nranges ? dense_base[nranges - 1].x + tok_count[nranges - 1].x : 0u;
    cnt = *n_extra_dev;
  }
  const uint32_t per = (cnt + slices - 1u) / slices;
  const uint32_t lo = min(cnt, sl * per), hi = min(cnt, lo + per);
  for (uint32_t i = lo + threadIdx.x; i < hi; i += 256u) {
    d_type[dst0 + i] = s_type[src0 + i];
    d_start[dst0 + i] = s_start[src0 + i];
    d_end[dst0 + i] = s_end[src0 + i];
    d_kw[dst0 + i] = s_kw[src0 + i];
  }
}
// ... and the statements' tok_begin moved along (thread per statement; its range by binary search of the cut table)
__global__ void __launch_bounds__(256) k_stmt_tok_remap(NutdbStmt* __restrict__ stmt, uint32_t nstmt, const uint32_t* __restrict__ cut_stmt,
                                                        const uint32_t* __restrict__ tok_base, const uint2* __restrict__ tok_count,
                                                        const uint2* __restrict__ dense_base, const uint32_t* __restrict__ nranges_dev) {
  const uint32_t s = blockIdx.x * 256u + threadIdx.x;
  const uint32_t nranges = *nranges_dev;
  if (s >= nstmt || nranges == 0u) return;
  const uint32_t tb = stmt[s].tok_begin;
  const uint32_t dense_total = dense_base[nranges - 1].x + tok_count[nranges - 1].x;
  if (tb >= tok_base[nranges]) {  // the exact lexer's region (and the "no tokens" value of empty statements)
    stmt[s].tok_begin = dense_total + (tb - tok_base[nranges]);
    return;
  }
  uint32_t lo = 0, hi = nranges;  // last range with cut_stmt[r] <= s
  while (hi - lo > 1u) {
    const uint32_t mid = (lo + hi) >> 1;
    if (cut_stmt[mid] <= s) lo = mid;
    else hi = mid;
  }
  stmt[s].tok_begin = dense_base[lo].x + (tb - tok_base[lo]);
}
