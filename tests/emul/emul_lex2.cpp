// TEST HARNESS (not shipped): the warp-cooperative lexer of nutdb_b200/csrc/lex2_core.cuh on the host.
// The per-lane / per-window functions are the device's own (NUTDB_HD); this file replaces the warp
// (ballots, shuffles) by loops over 32 lanes and the three kernels + scans by sequential passes,
// with the same decomposition: segments of `seg` bytes (one warp each on the device), windows of 32.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../nutdb_b200/csrc/lex_tables.hpp"
#include "../../nutdb_b200/csrc/lex2_core.cuh"

using namespace nlex;
using namespace nlex2;

namespace {
struct Env {
  const uint8_t* text;
  const uint32_t* bitmap;
  uint32_t n;
  const LexTables* T;
  const Lex2Tables* K;
  uint8_t byte(uint32_t p) const { return p < n ? text[p] : 0; }
  bool bnd(uint32_t p) const { return p < n && ((bitmap[p >> 5] >> (p & 31)) & 1u); }
};
struct HSrc {
  const Env& e;
  uint8_t byte(uint32_t p) const { return e.byte(p); }
  // mimic the device: contiguous only inside one 8 KB tile
  const uint8_t* span(uint32_t p, uint32_t len) const {
    return (p / 8192 == (p + len - 1) / 8192 && p + len <= e.n) ? e.text + p : nullptr;
  }
};

void entry_esc(const Env& e, uint32_t pos, uint8_t& prev, uint8_t& esc) {
  prev = 0;
  esc = 0;
  if (pos == 0) return;
  prev = e.byte(pos - 1);
  if (e.bnd(pos)) return;
  uint32_t nrun = 0, p = pos;
  while (p > 0 && e.byte(p - 1) == '\\') {
    nrun++;
    p--;
    if (e.bnd(p)) break;
  }
  esc = (uint8_t)(nrun & 1u);
}

struct WinSetup {
  Win w;
  Next nx;
  Events ev;
  uint32_t base, escm;
  uint8_t prev_byte;
};

// the device's setup_window for window `base` of the block starting at `blk`
void setup_window(const Env& e, uint32_t blk, uint32_t base, bool virt, WinSetup& u) {
  Win& w = u.w;
  std::memset(&w, 0, sizeof(w));
  u.base = base;
  for (int lane = 0; lane < 32; lane++) {
    uint32_t pos = base + lane;
    uint16_t k = pos < e.n ? e.K->cls[e.byte(pos)] : 0;
    uint32_t bit = 1u << lane;
    if (pos < e.n) w.valid |= bit;
    if (k & K_SQ) w.sq |= bit;
    if (k & K_DQ) w.dq |= bit;
    if (k & K_BT) w.bt |= bit;
    if (k & K_NL) w.nl |= bit;
    if (k & K_BS) w.bs |= bit;
    if (k & K_DASH) w.dash |= bit;
    if (k & K_SLASH) w.slash |= bit;
    if (k & K_STAR) w.star |= bit;
    if (k & K_L) w.L |= bit;
    if (k & K_D) w.D |= bit;
    if (k & K_DOT) w.DOT |= bit;
    if (k & K_OP) w.OP |= bit;
    if (k & K_P) w.P |= bit;
    if (k & K_WS) w.WS |= bit;
    if (k & K_IE) w.IE |= bit;
    if (k & K_NE) w.NE |= bit;
  }
  uint32_t bnd = base < e.n ? (e.bitmap[base >> 5] & w.valid) : 0u;
  if (virt && e.n >= base && e.n - base < 32u) bnd |= 1u << (e.n - base);
  w.bnd = bnd;
  uint32_t p = base + 32;
  if (p >= e.n) {
    u.nx.byte = 0;
    u.nx.bnd = 1;
    u.nx.cls = 0;
  } else {
    u.nx.byte = e.byte(p);
    u.nx.bnd = e.bnd(p) ? 1 : 0;
    u.nx.cls = e.K->cls[u.nx.byte];
  }
  uint8_t esc_in;
  if (base == blk) {  // lane 0 of the warp: look back through the text
    entry_esc(e, base < e.n ? base : 0u, u.prev_byte, esc_in);
  } else {            // other lanes: the previous window's backslash mask
    u.prev_byte = base <= e.n && base > 0 ? e.byte(base - 1) : 0;
    uint32_t bs_prev = 0;
    for (int lane = 0; lane < 32; lane++)
      if (base - 32 + lane < e.n && e.byte(base - 32 + lane) == '\\') bs_prev |= 1u << lane;
    int run = clz32(~bs_prev);
    esc_in = (uint8_t)(run >= 32 ? 0 : (run & 1));
  }
  u.escm = esc_mask32(w.bs, esc_in) & ~w.bnd;
  u.ev = make_events(w, u.escm, u.prev_byte);
}
uint32_t window_fn(const LexTables& T, const WinSetup& u) {
  if (u.ev.all) return ctx_window_fn(T, u.w, u.ev, NUTDB_VEC8_ID);
  return vec8_then_row(NUTDB_VEC8_ID, T.a_row[EV_OTHER][0], T.a_row[EV_OTHER][1]);
}

struct Out {
  uint8_t* type;
  uint32_t* start;
  uint32_t* end;
  uint8_t* kw;
  uint32_t cap;
  uint32_t* stmt_tok_begin;
  uint32_t* stmt_tok_end;
  const uint64_t* offs;
  uint64_t nstmt;
  std::vector<uint8_t>* flag;
  std::vector<uint32_t>* list;
  bool overflow = false;
  uint32_t find_stmt(uint32_t pos) const {  // last s with offs[s]-offs[0] <= pos
    uint64_t lo = 0, hi = nstmt;
    while (lo < hi) {
      uint64_t mid = (lo + hi) / 2;
      if ((uint32_t)(offs[mid] - offs[0]) > pos) hi = mid;
      else lo = mid + 1;
    }
    return (uint32_t)(lo - 1);
  }
  void punt(uint32_t pos) {
    uint32_t s = find_stmt(pos);
    if (!(*flag)[s]) {
      (*flag)[s] = 1;
      list->push_back(s);
    }
  }
  void token(uint32_t i, uint8_t t, uint32_t s, uint32_t e, uint8_t k) {
    if (i >= cap) { overflow = true; return; }
    type[i] = t;
    start[i] = s;
    end[i] = e;
    kw[i] = k;
  }
  // same O(1) statement lookup as the device sink (first_stmt table built by k_prep)
  const uint32_t* first_stmt = nullptr;
  uint32_t nbytes = 0;
  uint32_t off32(uint64_t s) const { return (uint32_t)(offs[s] - offs[0]); }
  void stmt_begin(uint32_t pos, uint32_t first) {
    uint32_t c = first_stmt[pos >> 5];
    while (off32(c) != pos || off32(c + 1) == pos) c++;
    stmt_tok_begin[c] = first;
    uint32_t p = c;
    while (p > 0) {
      p--;
      if (off32(p + 1) != off32(p)) {
        stmt_tok_end[p] = first;
        break;
      }
    }
  }
  void stmt_end(uint32_t pos, uint32_t endi) {
    if (pos + 1u != nbytes) return;
    uint64_t p = nstmt;
    while (p > 0) {
      p--;
      if (off32(p + 1) != off32(p)) {
        stmt_tok_end[p] = endi;
        break;
      }
    }
  }
};

struct StmtSrc {
  const uint8_t* text;
  uint32_t begin, end;
  uint8_t byte(uint32_t p) const { return p < end ? text[p] : 0; }
  bool boundary(uint32_t p) const { return p == begin; }
};
struct ExactSink {
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
LexTables g_T;
Lex2Tables g_K;
bool g_init = false;
uint64_t g_punts = 0;
std::vector<uint32_t> g_xoff;
}  // namespace

extern "C" {

// Sequential statement splitter: the context automaton of lex_core.cuh run byte by byte over the whole buffer
// (the definition the scan-based GPU splitter has to reproduce).  Returns n_stmt; offs gets n_stmt + 1 entries.
int64_t emul_split(const uint8_t* text, uint64_t n, uint64_t* offs, uint64_t cap) {
  uint64_t k = 0;
  offs[k++] = 0;
  uint8_t A = A_C, prev = 0, esc = 0;
  for (uint64_t pos = 0; pos < n; pos++) {
    const uint8_t b = text[pos];
    const uint8_t ev = a_event(b, prev, esc != 0);
    if (A <= A_CX && b == ';') {
      if (k >= cap) return -1;
      offs[k++] = pos + 1;
    }
    A = a_next(A, ev);
    esc = (b == '\\' && !esc) ? 1 : 0;
    prev = b;
  }
  bool tail = false;
  for (uint64_t pos = offs[k - 1]; pos < n; pos++)
    if (!(text[pos] == ' ' || text[pos] == '\t' || text[pos] == '\n' || text[pos] == '\r')) tail = true;
  if (tail) {
    if (k >= cap) return -1;
    offs[k++] = n;
  }
  return (int64_t)k - 1;
}

uint64_t emul_lex2_punts(void) { return g_punts; }

// Tokens of every statement (no Whitespace / Comment) + per-STATEMENT token ranges.
// Returns the number of tokens, -1 on capacity overflow, -2 if the scanned entry states disagree with a sequential walk.
int64_t emul_lex2(const uint8_t* text, uint32_t n, const uint64_t* offs, uint64_t nstmt, uint32_t seg_len,
                  uint8_t* tok_type, uint32_t* tok_start, uint32_t* tok_end, uint8_t* tok_kw, uint32_t cap,
                  uint32_t* stmt_tok_begin, uint32_t* stmt_tok_end) {
  if (!g_init) {
    build_lex_tables(g_T);
    build_lex2_tables(g_K);
    g_init = true;
  }
  if (n == 0) return 0;
  std::vector<uint32_t> bitmap((n + 31) / 32 + 2, 0);
  for (uint64_t s = 0; s < nstmt; s++)
    if (offs[s + 1] > offs[s]) {
      uint32_t p = (uint32_t)(offs[s] - offs[0]);
      bitmap[p >> 5] |= 1u << (p & 31);
    }
  Env e{text, bitmap.data(), n, &g_T, &g_K};
  const uint32_t blk_len = seg_len;  // bytes per warp on the device (32 windows = 1024 there)
  const uint32_t nblk = (n + blk_len - 1) / blk_len, nwin = (n + 31) / 32;
  // phase 1: per-window transition functions -> per-block functions -> block entry states
  std::vector<uint32_t> wfn(nwin), bfn(nblk);
  for (uint32_t g = 0; g < nblk; g++) {
    uint32_t run = NUTDB_VEC8_ID;
    for (uint32_t base = g * blk_len; base < std::min((g + 1) * blk_len, n); base += 32) {
      WinSetup u;
      setup_window(e, g * blk_len, base, false, u);
      wfn[base >> 5] = window_fn(g_T, u);
      run = vec8_then(run, wfn[base >> 5]);
    }
    bfn[g] = run;
  }
  std::vector<uint8_t> entA(nblk);
  uint32_t pref = NUTDB_VEC8_ID;
  for (uint32_t g = 0; g < nblk; g++) {
    entA[g] = (uint8_t)vec8_apply(pref, A_C);
    pref = vec8_then(pref, bfn[g]);
  }
  // phase 2 (count + flags) and phase 3 (emit): identical per-window work, carries passed along
  std::vector<uint8_t> flag(nstmt + 1, 0);
  std::vector<uint32_t> list;
  Out out{tok_type, tok_start, tok_end, tok_kw, 0, stmt_tok_begin, stmt_tok_end, offs, nstmt, &flag, &list};
  std::vector<uint32_t> first_stmt((n + 31) / 32 + 2, 0xFFFFFFFFu);
  for (uint64_t st = 0; st < nstmt; st++)
    if (offs[st + 1] > offs[st]) {
      uint32_t p = (uint32_t)(offs[st] - offs[0]);
      first_stmt[p >> 5] = std::min(first_stmt[p >> 5], (uint32_t)st);
    }
  out.first_stmt = first_stmt.data();
  out.nbytes = n;
  HSrc src{e};
  std::vector<uint32_t> whas(nwin);
  uint32_t ntok_main = 0;
  for (int pass = 0; pass < 2; pass++) {
    uint32_t count = 0, stmt_start = 0;
    StrCarry sc;
    Hist hprev;
    if (pass == 1) {
      if ((uint64_t)ntok_main > cap) return -1;
      out.cap = ntok_main;
    }
    for (uint32_t g = 0; g < nblk; g++) {
      uint8_t s = entA[g];
      const uint32_t blk = g * blk_len;
      for (uint32_t base = blk; base < std::min(blk + blk_len, n); base += 32) {
        WinSetup u;
        setup_window(e, blk, base, true, u);
        WinCtx o;
        ctx_window(u.w, u.ev, base, u.nx, s, u.prev_byte, o);
        o.escm = u.escm;
        // the symbolic function (phase 1) and the concrete walk must agree
        // (not in the window holding the batch end: only the concrete walk sees the virtual statement start there)
        if (base + 32 <= n && (uint8_t)vec8_apply(wfn[base >> 5], s) != o.s_out) return -2;
        Hist h;
        if (base == blk) {  // lane 0: raw classes of the 32 bytes in front of the block if it is entered in code
          if (blk >= 32) {
            if (entA[g] <= A_CX)
              for (int lane = 0; lane < 32; lane++) {
                uint16_t k = e.K->cls[e.byte(blk - 32 + lane)];
                uint32_t bit = 1u << lane;
                if (k & K_L) h.L |= bit;
                if (k & K_D) h.D |= bit;
                if (k & K_DOT) h.DOT |= bit;
                if (k & K_OP) h.OP |= bit;
              }
            h.bnd = e.bitmap[(blk - 32) >> 5];
          }
        } else {
          h = hprev;
        }
        uint32_t bad = 0;
        if (pass == 0) {
          StrCarry none;
          (void)none;
          const uint32_t has = win_has_mask(g_T, src, u.w, o, h, u.nx, base, u.prev_byte, bad);
          const uint32_t nt = (uint32_t)popc32(has) + (uint32_t)popc32(win_eof_mask(u.w, u.nx));
          if (u.w.bs == 0xFFFFFFFFu) bad |= 1u;
          whas[base >> 5] = has;
          for (int i = 0; i < 32; i++) {
            if ((bad >> i) & 1u) out.punt(base + i);
            if ((o.bad_prev >> i) & 1u) out.punt(base + i - 1);
          }
          if (base + 32 == n && (o.s_out == A_SQ || o.s_out == A_DQ || o.s_out == A_BT || o.s_out == A_BC0 || o.s_out == A_BC))
            out.punt(n - 1);
          count += nt;
        } else {
          const uint32_t has = whas[base >> 5];
          // the device's emitting pass does not walk the events again: it unpacks what the counting pass stored
          WinCtx o2;
          ctx_unpack(ctx_pack(o, base), base, o2);
          o2.escm = o.escm;
          win_emit(g_T, src, out, u.w, o2, h, u.nx, base, u.prev_byte, sc, stmt_start, count, has);
          count += (uint32_t)popc32(has) + (uint32_t)popc32(win_eof_mask(u.w, u.nx));
        }
        sc = str_then(sc, o.sc);
        if (o.last_bnd1) stmt_start = o.last_bnd1 - 1;
        hprev.L = u.w.L & o.ct;
        hprev.D = u.w.D & o.ct;
        hprev.DOT = u.w.DOT & o.ct;
        hprev.OP = u.w.OP & o.ct;
        hprev.bnd = u.w.bnd;
        s = o.s_out;
      }
    }
    if (pass == 0) {
      ntok_main = count;
      // exact path: counts of flagged statements (needed before the token arrays can be sized)
      g_xoff.assign(list.size() + 1, 0);
      for (size_t i = 0; i < list.size(); i++) {
        uint32_t st = list[i];
        StmtSrc ssrc{text, (uint32_t)(offs[st] - offs[0]), (uint32_t)(offs[st + 1] - offs[0])};
        ExactSink sink{nullptr, nullptr, nullptr, nullptr, 0};
        LexCarry c;
        c.stmt_start = ssrc.begin;
        c.tok_start = ssrc.begin;
        Walker<false, StmtSrc, ExactSink> w(g_T, ssrc, sink, c);
        w.counting = true;
        for (uint32_t pos = ssrc.begin; pos < ssrc.end; pos++) w.step(pos, ssrc.byte(pos), pos == ssrc.begin, true);
        w.flush_eof(ssrc.end);
        g_xoff[i + 1] = g_xoff[i] + w.c.count;
      }
      if ((uint64_t)ntok_main + g_xoff[list.size()] > cap) return -1;
    }
  }
  if (out.overflow) return -1;
  std::vector<uint32_t>& xoff = g_xoff;
  uint64_t ntok = (uint64_t)ntok_main + xoff[list.size()];
  g_punts += list.size();
  for (size_t i = 0; i < list.size(); i++) {
    uint32_t s = list[i];
    StmtSrc src{text, (uint32_t)(offs[s] - offs[0]), (uint32_t)(offs[s + 1] - offs[0])};
    ExactSink sink{tok_type, tok_start, tok_end, tok_kw, (uint32_t)ntok};
    LexCarry c;
    c.stmt_start = src.begin;
    c.tok_start = src.begin;
    c.count = ntok_main + xoff[i];
    Walker<false, StmtSrc, ExactSink> w(g_T, src, sink, c);
    for (uint32_t pos = src.begin; pos < src.end; pos++) w.step(pos, src.byte(pos), pos == src.begin, true);
    w.flush_eof(src.end);
    stmt_tok_begin[s] = ntok_main + xoff[i];
    stmt_tok_end[s] = w.c.count;
  }
  return (int64_t)ntok;
}

}  // extern "C"
