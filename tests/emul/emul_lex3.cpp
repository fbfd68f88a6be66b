// TEST HARNESS (not shipped): the single-pass lexer of nutdb_b200/csrc/lex3_core.cuh on the host.
// The per-window / per-token functions are the device's own (NUTDB_HD); this file replaces the kernel's
// plumbing (tiles, warp shuffles, the two look-back scans) by one sequential loop over the windows that
// passes the carries along, with the same decomposition: blocks of `seg` bytes (one warp each on the
// device, whose lane 0 sees its predecessor differently from lanes 1..31), windows of 32 bytes.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../nutdb_b200/csrc/lex_tables.hpp"
#include "../../nutdb_b200/csrc/lex3_core.cuh"

using namespace nlex;
using namespace nlex2;
using namespace nlex3;

namespace {
struct Env3 {
  const uint8_t* text;
  const uint32_t* bitmap;
  uint32_t n;
  uint8_t byte(uint32_t p) const { return p < n ? text[p] : 0; }
  bool bnd(uint32_t p) const { return p < n && ((bitmap[p >> 5] >> (p & 31)) & 1u); }
};
struct HSrc3 {
  const Env3& e;
  uint8_t byte(uint32_t p) const { return e.byte(p); }
  uint8_t at(uint32_t p) const { return e.byte(p); }
  const uint8_t* span(uint32_t p, uint32_t len) const {
    return (p / 8192 == (p + len - 1) / 8192 && p + len <= e.n) ? e.text + p : nullptr;  // like the device: inside one tile
  }
};
struct StmtSrc3 {
  const uint8_t* text;
  uint32_t begin, end;
  uint8_t byte(uint32_t p) const { return p < end ? text[p] : 0; }
  bool boundary(uint32_t p) const { return p == begin; }
};
struct ExactSink3 {
  uint8_t* type;
  uint32_t* start;
  uint32_t* end;
  uint8_t* kw;
  uint32_t cap;
  bool overflow = false;
  void token(uint32_t i, uint8_t t, uint32_t s, uint32_t e, uint8_t k) {
    if (i >= cap) { overflow = true; return; }
    type[i] = t;
    start[i] = s;
    end[i] = e;
    kw[i] = k;
  }
  void seg_begin(uint32_t, uint32_t, uint32_t) {}
  void seg_end(uint32_t, uint32_t, uint32_t) {}
};
struct Rec3 {
  uint32_t start, end, flags;
};
LexTables g_T3;
bool g_init3 = false;
uint64_t g_punts3 = 0;
std::vector<uint32_t> g_dbg3;  // per 8 KB tile: entry state, first token, statement start, open quote
}  // namespace

extern "C" {

uint64_t emul_lex3_punts(void) { return g_punts3; }
uint32_t emul_lex3_debug_tiles(uint32_t* out, uint32_t cap_tiles) {
  const uint32_t nt = std::min<uint32_t>(cap_tiles, (uint32_t)(g_dbg3.size() / 4));
  std::memcpy(out, g_dbg3.data(), 16 * (size_t)nt);
  return nt;
}

// self-check of the bit-slicing: class masks of 32 arbitrary bytes against a per-byte definition.  Returns 0 if equal.
int emul_lex3_check_classes(const uint8_t* bytes32) {
  uint32_t v[8], p[8];
  std::memcpy(v, bytes32, 32);
  bit_planes(v, p);
  for (int k = 0; k < 8; k++)
    for (int i = 0; i < 32; i++)
      if (((p[k] >> i) & 1u) != ((bytes32[i] >> k) & 1u)) return 1;
  Win w;
  Ops op;
  classify_planes(p, 0xFFFFFFFFu, w, op);
  for (int i = 0; i < 32; i++) {
    const uint8_t c = bytes32[i];
    const uint32_t bit = 1u << i;
    auto is = [&](uint32_t m) { return (m & bit) != 0; };
    const bool L = (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_', D = c >= '0' && c <= '9';
    const bool P = c && std::strchr("()[]{},:+%&|^~;*", c) != nullptr;
    if (is(w.L) != L || is(w.D) != D || is(w.P) != P) return 2;
    if (is(w.sq) != (c == '\'') || is(w.dq) != (c == '"') || is(w.bt) != (c == '`') || is(w.bs) != (c == '\\')) return 3;
    if (is(w.nl) != (c == '\n' || c == '\r') || is(w.WS) != (c == '\n' || c == '\r' || c == ' ' || c == '\t')) return 4;
    if (is(w.dash) != (c == '-') || is(w.slash) != (c == '/') || is(w.star) != (c == '*') || is(w.DOT) != (c == '.')) return 5;
    if (is(op.LT) != (c == '<') || is(op.GT) != (c == '>') || is(op.EQ) != (c == '=') || is(op.BANG) != (c == '!')) return 6;
    if (is(w.OP) != (c == '<' || c == '>' || c == '=' || c == '!')) return 7;
  }
  return 0;
}

// Tokens of every statement (no Whitespace / Comment) + per-STATEMENT token ranges.
// Returns the number of tokens, -1 on capacity overflow, -2 if the scanned entry states disagree with a sequential walk.
int64_t emul_lex3(const uint8_t* text, uint32_t n, const uint64_t* offs, uint64_t nstmt, uint32_t seg_len, uint8_t* tok_type,
                  uint32_t* tok_start, uint32_t* tok_end, uint8_t* tok_kw, uint32_t cap, uint32_t* stmt_tok_begin,
                  uint32_t* stmt_tok_end) {
  if (!g_init3) {
    build_lex_tables(g_T3);
    g_init3 = true;
  }
  if (n == 0) return 0;
  const LexTables& T = g_T3;
  std::vector<uint32_t> bitmap((n + 31) / 32 + 8, 0);
  std::vector<uint32_t> off32(nstmt + 1);
  for (uint64_t s = 0; s <= nstmt; s++) off32[s] = (uint32_t)(offs[s] - offs[0]);
  for (uint64_t s = 0; s < nstmt; s++)
    if (off32[s + 1] > off32[s]) bitmap[off32[s] >> 5] |= 1u << (off32[s] & 31);
  std::vector<uint32_t> first_stmt((n + 31) / 32 + 8, 0xFFFFFFFFu);
  for (uint64_t s = 0; s < nstmt; s++)
    if (off32[s + 1] > off32[s]) first_stmt[off32[s] >> 5] = std::min(first_stmt[off32[s] >> 5], (uint32_t)s);
  Env3 e{text, bitmap.data(), n};
  HSrc3 src{e};
  std::vector<uint8_t> flag(nstmt + 1, 0);
  auto punt_stmt_at = [&](uint32_t sst) {
    uint32_t c = first_stmt[sst >> 5];
    while (off32[c] != sst || off32[c + 1] == sst) c++;
    flag[c] = 1;
  };
  const uint32_t nwin = (n + 31) / 32 + 1;  // (+ the window that starts exactly at the batch end, if any)
  std::vector<uint32_t> win_idx(nwin + 1, 0), win_has(nwin + 1, 0), win_eof(nwin + 1, 0), win_bnd(nwin + 2, 0), sst_in(nwin + 1, 0);
  std::vector<Rec3> recs;
  // carries
  uint8_t s_state = A_C, s_block = A_C;
  uint32_t count = 0, stmt_start1 = 0;  // 1 + last statement start
  StrCarry sc;
  Hist3 hprev;
  uint32_t bs_prev = 0;
  g_dbg3.clear();
  for (uint32_t base = 0; base <= n; base += 32) {
    const bool live = base < n;
    if (live && base % 8192 == 0) {
      g_dbg3.push_back(s_state);
      g_dbg3.push_back(count);
      g_dbg3.push_back(stmt_start1);
      g_dbg3.push_back(sc.open_pos | ((uint32_t)(sc.has_open != 0) << 30) | ((uint32_t)(sc.esc != 0) << 31));
    }
    const uint32_t blk = base - base % seg_len;
    if (base == blk) s_block = s_state;
    Win w;
    Ops op;
    std::memset(&w, 0, sizeof(w));
    const uint32_t valid = base + 32u <= n ? 0xFFFFFFFFu : (n > base ? ((1u << (n - base)) - 1u) : 0u);
    {
      uint8_t bytes[32];
      for (int i = 0; i < 32; i++) bytes[i] = (uint8_t)(base + i < n + 16 ? (base + i < n ? text[base + i] : 0x5A) : 0);  // junk behind the end
      uint32_t v[8], p[8];
      std::memcpy(v, bytes, 32);
      bit_planes(v, p);
      classify_planes(p, valid, w, op);
    }
    uint32_t bnd = live ? (bitmap[base >> 5] & valid) : 0u;
    if (n >= base && n - base < 32u) bnd |= 1u << (n - base);
    w.bnd = bnd;
    win_bnd[base >> 5] = bnd;
    Next nx;
    if (base + 32 >= n) {
      nx.byte = 0;
      nx.bnd = 1;
      nx.cls = 0;
    } else {
      nx.byte = e.byte(base + 32);
      nx.bnd = e.bnd(base + 32) ? 1 : 0;
      const uint8_t pr = T.prop[nx.byte];
      nx.cls = (uint16_t)(((pr & PR_IDENT_END) ? 1u : 0u) | ((pr & PR_NUM_END) ? 2u : 0u));
    }
    uint8_t prev_byte = 0, prev2_byte = 0, esc_in = 0;
    if (base == blk) {
      if (base > 0 && base <= n) {
        prev_byte = e.byte(base - 1);
        if (!(w.bnd & 1u)) {
          uint32_t nrun = 0, p = base;
          while (p > 0 && e.byte(p - 1) == '\\') {
            nrun++;
            p--;
            if (e.bnd(p)) break;
          }
          esc_in = (uint8_t)(nrun & 1u);
        }
      }
    } else {
      prev_byte = base <= n && base > 0 ? e.byte(base - 1) : 0;
      const int run = clz32(~bs_prev);
      esc_in = (uint8_t)(run >= 32 ? 0 : (run & 1));
    }
    prev2_byte = base <= n && base > 1 ? e.byte(base - 2) : 0;
    const uint32_t escm = esc_mask32(w.bs, esc_in) & ~w.bnd;
    const Events ev = make_events(w, escm, prev_byte);
    uint32_t fnv = NUTDB_VEC8_ID;
    if (live)
      fnv = ev.all ? ctx_window_fn(T, w, ev, NUTDB_VEC8_ID) : vec8_then_row(NUTDB_VEC8_ID, T.a_row[EV_OTHER][0], T.a_row[EV_OTHER][1]);
    WinCtx3 o;
    if (live) ctx_window3(w, ev, base, nx, s_state, prev_byte, o);
    o.escm = escm;
    if (live && base + 32 <= n && (uint8_t)vec8_apply(fnv, s_state) != o.s_out) return -2;
    Hist3 h;
    if (base == blk) {
      if (blk >= 32 && s_block <= A_CX)
        for (int lane = 0; lane < 32; lane++) {
          const uint8_t c = e.byte(blk - 32 + lane);
          const uint8_t pr = T.prop[c];
          if ((pr & PR_WORD) && !(pr & PR_DIGIT)) h.L |= 1u << lane;
          if (pr & PR_DIGIT) h.D |= 1u << lane;
          if (c == '.') h.DOT |= 1u << lane;
        }
      h.bnd = blk >= 32 ? bitmap[(blk - 32) >> 5] : 0u;
    } else {
      h = hprev;
    }
    TokMasks m;
    const uint32_t sst_open = stmt_start1 ? stmt_start1 - 1 : 0;
    sst_in[base >> 5] = sst_open;
    if (live) {
      win_tokens3(w, op, o, h, nx, prev_byte, prev2_byte, m);
      uint32_t bad = m.bad | win_bad_mask(src, w, o, base, prev_byte);
      if (w.bs == 0xFFFFFFFFu) bad |= 1u;
      const uint32_t bnds = w.bnd & w.valid;
      for (int i = 0; i < 32; i++) {
        if ((bad >> i) & 1u) {
          const uint32_t below = bnds & (i >= 31 ? 0xFFFFFFFFu : ((2u << i) - 1u));
          punt_stmt_at(below ? base + (uint32_t)(31 - clz32(below)) : sst_open);
        }
        if ((o.bad_prev >> i) & 1u) {
          const uint32_t below = bnds & ((1u << i) - 1u);
          punt_stmt_at(below ? base + (uint32_t)(31 - clz32(below)) : sst_open);
        }
      }
      if (base + 32 == n && (o.s_out == A_SQ || o.s_out == A_DQ || o.s_out == A_BT || o.s_out == A_BC0 || o.s_out == A_BC))
        punt_stmt_at(bnds ? base + (uint32_t)(31 - clz32(bnds)) : sst_open);
    }
    win_idx[base >> 5] = count;
    win_has[base >> 5] = m.has;
    win_eof[base >> 5] = m.eofm;
    if (live) {
      auto rec = [&](uint32_t idx, uint32_t start_abs, uint32_t end_abs, uint32_t flags) {
        if (recs.size() <= idx) recs.resize(idx + 1);
        recs[idx] = Rec3{start_abs, end_abs, flags};
      };
      win_records3(o, m, base, sc, count, rec);
      count += (uint32_t)popc32(m.has) + (uint32_t)popc32(m.eofm);
      sc = str_then(sc, o.sc);
      if (o.last_bnd1) stmt_start1 = o.last_bnd1;
      hprev.L = w.L & o.ct;
      hprev.D = w.D & o.ct;
      hprev.DOT = w.DOT & o.ct;
      hprev.bnd = w.bnd;
      bs_prev = w.bs;
      s_state = o.s_out;
    }
  }
  const uint32_t ntok_main = count;
  if (recs.size() != ntok_main) return -3;
  if (ntok_main > cap) return -1;
  // thread per token
  for (uint32_t k = 0; k < ntok_main; k++) {
    const Rec3& r = recs[k];
    const uint32_t last = r.end - 1u, wv = last >> 5, i = last & 31u;
    const uint32_t bb = win_bnd[wv] & (i >= 31u ? 0xFFFFFFFFu : ((2u << i) - 1u));
    const uint32_t sst = bb ? 32u * wv + (uint32_t)(31 - clz32(bb)) : sst_in[wv];
    Tok3 tk;
    if (token_finish3(T, src, r.start, r.end, r.flags, sst, tk)) tk.kw = token_keyword3(T, src, r.start, r.end - r.start);
    if (tk.punt) {
      punt_stmt_at(sst);
      tk.type = NUTDB_TT_POISON;
      tk.start = tk.end = 0;
      tk.kw = 0;
    }
    tok_type[k] = tk.type;
    tok_start[k] = tk.start;
    tok_end[k] = tk.end;
    tok_kw[k] = tk.kw;
  }
  // statement token ranges: the parser kernel's formula (StmtToks::index_at)
  auto index_at = [&](uint32_t pos) {
    const uint32_t wv = pos >> 5, b = pos & 31u;
    uint32_t r = win_idx[wv];
    if (b) {
      const uint32_t below = (1u << b) - 1u;
      r += (uint32_t)popc32(win_has[wv] & below) + (uint32_t)popc32(win_eof[wv] & below);
    }
    return r;
  };
  for (uint64_t s = 0; s < nstmt; s++) {
    if (off32[s + 1] == off32[s]) continue;
    stmt_tok_begin[s] = index_at(off32[s]);
    stmt_tok_end[s] = index_at(off32[s + 1]);
  }
  // exact path for the flagged statements, in statement order, into the region behind the main tokens
  uint64_t ntok = ntok_main;
  for (uint64_t s = 0; s < nstmt; s++) {
    if (!flag[s]) continue;
    g_punts3++;
    StmtSrc3 ssrc{text, off32[s], off32[s + 1]};
    ExactSink3 sink{tok_type, tok_start, tok_end, tok_kw, cap};
    LexCarry c;
    c.stmt_start = ssrc.begin;
    c.tok_start = ssrc.begin;
    c.count = (uint32_t)ntok;
    Walker<false, StmtSrc3, ExactSink3> wk(T, ssrc, sink, c);
    for (uint32_t pos = ssrc.begin; pos < ssrc.end; pos++) wk.step(pos, ssrc.byte(pos), pos == ssrc.begin, true);
    wk.flush_eof(ssrc.end);
    if (sink.overflow) return -1;
    stmt_tok_begin[s] = (uint32_t)ntok;
    stmt_tok_end[s] = wk.c.count;
    ntok = wk.c.count;
  }
  return (int64_t)ntok;
}

}  // extern "C"
