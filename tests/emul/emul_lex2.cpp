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
};

void build_win(const Env& e, uint32_t base, uint8_t* b, uint16_t* k, Win& w) {
  std::memset(&w, 0, sizeof(w));
  for (int lane = 0; lane < 32; lane++) {
    uint32_t pos = base + lane;
    b[lane] = e.byte(pos);
    k[lane] = pos < e.n ? e.K->cls[b[lane]] : 0;
    uint32_t bit = 1u << lane;
    if (pos < e.n) w.valid |= bit;
    if (k[lane] & K_SQ) w.sq |= bit;
    if (k[lane] & K_DQ) w.dq |= bit;
    if (k[lane] & K_BT) w.bt |= bit;
    if (k[lane] & K_NL) w.nl |= bit;
    if (k[lane] & K_BS) w.bs |= bit;
    if (k[lane] & K_DASH) w.dash |= bit;
    if (k[lane] & K_SLASH) w.slash |= bit;
    if (k[lane] & K_STAR) w.star |= bit;
    if (k[lane] & K_L) w.L |= bit;
    if (k[lane] & K_D) w.D |= bit;
    if (k[lane] & K_DOT) w.DOT |= bit;
    if (k[lane] & K_OP) w.OP |= bit;
  }
  uint32_t bnd = e.bitmap[base >> 5] & w.valid;
  if (e.n >= base && e.n - base < 32u) bnd |= 1u << (e.n - base);
  w.bnd = bnd;
}
Next next_of(const Env& e, uint32_t base) {
  Next nx;
  uint32_t p = base + 32;
  if (p >= e.n) {
    nx.byte = 0;
    nx.bnd = 1;
    nx.cls = 0;
  } else {
    nx.byte = e.byte(p);
    nx.bnd = e.bnd(p) ? 1 : 0;
    nx.cls = e.K->cls[nx.byte];
  }
  return nx;
}
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
uint32_t esc_ballot(const Win& w, uint8_t carry) {
  uint32_t m = 0;
  for (int lane = 0; lane < 32; lane++)
    if (lane_esc(w.bs, lane, carry)) m |= 1u << lane;
  return m & ~w.bnd;
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
};

// one segment; Emit=false: returns its CSum and flags statements; Emit=true: writes tokens
template <bool Emit>
CSum walk_segment(const Env& e, uint32_t seg, uint32_t seg_len, uint8_t entryA, const CSum& pre, Out& out) {
  Carry2 c;
  bool opened = false;
  uint32_t nbnd_seen = 0;
  c.s = entryA;
  entry_esc(e, seg, c.prev, c.esc);
  c.str_start = NUTDB_NO_TOK;
  c.stmt_start = 0;
  if (Emit) {
    c.count = pre.count;
    c.str_start = pre.tok_start;
    c.escaped = pre.escaped;
    c.stmt_start = pre.stmt_start;
  }
  Hist h;
  if (seg >= 32) {
    if (c.s <= A_CX)
      for (int lane = 0; lane < 32; lane++) {
        uint16_t k = e.K->cls[e.byte(seg - 32 + lane)];
        uint32_t bit = 1u << lane;
        if (k & K_L) h.L |= bit;
        if (k & K_D) h.D |= bit;
        if (k & K_DOT) h.DOT |= bit;
        if (k & K_OP) h.OP |= bit;
      }
    h.bnd = e.bitmap[(seg - 32) >> 5];
  }
  {
    uint8_t b0 = e.byte(seg);
    bool bnd0 = e.bnd(seg);
    c.reopen = (c.s == A_C && !bnd0 && (c.prev == '\'' || c.prev == '"') && b0 == c.prev) ? 1 : 0;
  }
  HSrc src{e};
  uint32_t end = std::min(seg + seg_len, e.n);
  for (uint32_t base = seg; base < end; base += 32) {
    uint8_t b[32];
    uint16_t k[32];
    Win w;
    build_win(e, base, b, k, w);
    Next nx = next_of(e, base);
    uint32_t escm = esc_ballot(w, c.esc);
    Events ev = make_events(w, escm, c.prev);
    uint32_t stmt_entry = c.stmt_start, str_before = c.str_start;
    // stage 2 is warp-uniform: every lane runs it on identical inputs; only the capture differs
    CtxOut outs[32];
    Carry2 cnext = c;
    for (int lane = 0; lane < 32; lane++) {
      Carry2 cc = c;
      ctx_window(w, ev, base, nx, lane, cc, outs[lane]);
      if (lane == 0) cnext = cc;
    }
    const uint8_t prev_byte = c.prev;
    c = cnext;
    if (c.str_start != str_before) opened = true;
    nbnd_seen += (uint32_t)popc32(w.bnd & w.valid);
    LaneTok t[32];
    uint32_t tokmask = 0, eofmask = 0, badmask = 0;
    for (int lane = 0; lane < 32; lane++) {
      t[lane] = lane_token(*e.T, src, lane, base, b[lane], k[lane], w, outs[lane], h, nx, escm, prev_byte);
      if (t[lane].has) tokmask |= 1u << lane;
      if (t[lane].eof) eofmask |= 1u << lane;
      if (t[lane].bad) badmask |= 1u << lane;
    }
    const CtxOut& o = outs[0];
    if (!Emit) {
      badmask |= o.bad;
      for (int lane = 0; lane < 32; lane++) {
        if ((badmask >> lane) & 1u) out.punt(base + lane);
        if ((o.bad_prev >> lane) & 1u) out.punt(base + lane - 1);
      }
    } else {
      for (int lane = 0; lane < 32; lane++) {
        uint32_t lt = (1u << lane) - 1u;
        uint32_t idx = c.count + (uint32_t)popc32(tokmask & lt) + (uint32_t)popc32(eofmask & lt);
        uint32_t below = w.bnd & w.valid & (lt | (1u << lane));
        uint32_t sst = below ? base + (uint32_t)(31 - clz32(below)) : stmt_entry;
        if ((w.bnd & w.valid) & (1u << lane)) out.stmt_tok_begin[out.find_stmt(base + lane)] = idx;
        if (t[lane].has) {
          if (idx < out.cap) {
            out.type[idx] = t[lane].type;
            out.start[idx] = t[lane].start - sst;
            out.end[idx] = t[lane].end - sst;
            out.kw[idx] = t[lane].kw;
          } else {
            out.overflow = true;
          }
        }
        if (t[lane].eof) {
          uint32_t ei = idx + t[lane].has;
          if (ei < out.cap) {
            out.type[ei] = NUTDB_TT_EOF;
            out.start[ei] = base + lane + 1 - sst;
            out.end[ei] = base + lane + 1 - sst;
            out.kw[ei] = 0;
          } else {
            out.overflow = true;
          }
          out.stmt_tok_end[out.find_stmt(base + lane)] = ei + 1;
        }
      }
    }
    c.count += (uint32_t)popc32(tokmask) + (uint32_t)popc32(eofmask);
    h.L = w.L & o.ct;
    h.D = w.D & o.ct;
    h.DOT = w.DOT & o.ct;
    h.OP = w.OP & o.ct;
    h.bnd = w.bnd;
    c.esc = esc_carry_out(w.bs, c.esc);
    c.prev = b[31];
  }
  if (!Emit && end == e.n && (e.n & 31u) == 0u &&
      (c.s == A_SQ || c.s == A_DQ || c.s == A_BT || c.s == A_BC0 || c.s == A_BC))
    out.punt(e.n - 1);
  CSum s;
  s.count = c.count;
  s.nseg = nbnd_seen;
  s.has_tok = opened ? 1 : 0;
  s.tok_start = c.str_start;
  s.escaped = c.escaped;
  s.stmt_start = c.stmt_start;
  return s;
}

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
}  // namespace

extern "C" {

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
  uint32_t nseg = (n + seg_len - 1) / seg_len;
  // phase 1: per-segment transition functions, then entry states
  std::vector<uint32_t> fn(nseg);
  for (uint32_t g = 0; g < nseg; g++) {
    uint32_t seg = g * seg_len, end = std::min(seg + seg_len, n);
    uint8_t prev, esc;
    entry_esc(e, seg, prev, esc);
    uint32_t run = NUTDB_VEC8_ID;
    for (uint32_t base = seg; base < end; base += 32) {
      uint8_t b[32];
      uint16_t k[32];
      Win w;
      build_win(e, base, b, k, w);
      w.bnd = e.bitmap[base >> 5] & w.valid;  // the function kernel does not add the virtual end
      w.L = w.D = w.DOT = w.OP = 0;
      uint32_t escm = esc_ballot(w, esc);
      Events ev = make_events(w, escm, prev);
      if (ev.all) run = ctx_window_fn(g_T, w, ev, run);
      else run = vec8_then_row(run, g_T.a_row[EV_OTHER][0], g_T.a_row[EV_OTHER][1]);
      esc = esc_carry_out(w.bs, esc);
      prev = b[31];
    }
    fn[g] = run;
  }
  std::vector<uint8_t> entA(nseg);
  uint32_t pref = NUTDB_VEC8_ID;
  for (uint32_t g = 0; g < nseg; g++) {
    entA[g] = (uint8_t)vec8_apply(pref, A_C);
    pref = vec8_then(pref, fn[g]);
  }
  // phase 2: counts + flags
  std::vector<uint8_t> flag(nstmt + 1, 0);
  std::vector<uint32_t> list;
  Out out{tok_type, tok_start, tok_end, tok_kw, 0, stmt_tok_begin, stmt_tok_end, offs, nstmt, &flag, &list};
  std::vector<CSum> sums(nseg), pre(nseg);
  CSum zero = csum_identity();
  for (uint32_t g = 0; g < nseg; g++) sums[g] = walk_segment<false>(e, g * seg_len, seg_len, entA[g], zero, out);
  CSum run = csum_identity();
  for (uint32_t g = 0; g < nseg; g++) {
    pre[g] = run;
    run = csum_then(run, sums[g]);
  }
  uint32_t ntok_main = run.count;
  // consistency of the scanned context states with a plain sequential walk (checks ctx_window_fn against ctx_window)
  {
    uint8_t s = A_C;
    for (uint32_t g = 0; g < nseg; g++) {
      if (s != entA[g]) return -2;
      s = (uint8_t)vec8_apply(fn[g], s);
    }
  }
  g_punts += list.size();
  // exact path: counts of flagged statements
  std::vector<uint32_t> xoff(list.size() + 1, 0);
  for (size_t i = 0; i < list.size(); i++) {
    uint32_t s = list[i];
    StmtSrc src{text, (uint32_t)(offs[s] - offs[0]), (uint32_t)(offs[s + 1] - offs[0])};
    ExactSink sink{nullptr, nullptr, nullptr, nullptr, 0};
    LexCarry c;
    c.stmt_start = src.begin;
    c.tok_start = src.begin;
    Walker<false, StmtSrc, ExactSink> w(g_T, src, sink, c);
    w.counting = true;
    for (uint32_t pos = src.begin; pos < src.end; pos++) w.step(pos, src.byte(pos), pos == src.begin, true);
    w.flush_eof(src.end);
    xoff[i + 1] = xoff[i] + w.c.count;
  }
  uint64_t ntok = (uint64_t)ntok_main + xoff[list.size()];
  if (ntok > cap) return -1;
  // phase 3: emit
  out.cap = ntok_main;
  for (uint32_t g = 0; g < nseg; g++) walk_segment<true>(e, g * seg_len, seg_len, entA[g], pre[g], out);
  if (out.overflow) return -1;
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
