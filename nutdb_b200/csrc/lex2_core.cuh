// lex2_core.cuh -- the mask-based lexer: one THREAD per 32-byte window, masks built by the warp.
//
// lex_core.cuh walks every byte through a branchy state machine; exact for every input, but ~60
// warp-instructions per byte on the GPU.  Here a warp takes 32 consecutive windows (1 KB):
//
//   stage 1 (warp)      for each window the 32 lanes classify one byte each and 14 ballots give the
//                       window's class masks; lane j keeps the masks of window j (a bit transposition)
//   stage 2 (thread)    the CONTEXT automaton (strings, quoted identifiers, comments) only steps at its
//                       sparse event bytes (quotes, newlines, "--", "/*", "*/", statement starts): the
//                       thread iterates over the set bits of its event mask -- first symbolically (a
//                       transition FUNCTION over the 8 states, so entry states come from a scan), then
//                       concretely; string / quoted-identifier tokens are found here at their closing quote
//   stage 3 (thread)    code tokens: for every byte where a token can END the thread finds the start by
//                       bit scanning its own and the previous window's masks (tokens longer than that
//                       look-back go to the exact path), applies the reference's end-of-token rule and
//                       looks up keywords
//   carries (warp scan) entry state, open string (offset, escaped flag), statement start, token count
//
// Only VALID text is handled natively.  Whatever the reference would reject with a lex error
// (tokenizer/mod.rs error sites) and a few rare valid forms (hex literals, `$n`, `@name`, runs of
// three or more of `< > = !`, code tokens longer than the look-back, more than 4 literals closing in
// one window) mark their STATEMENT for an exact re-lex by the walker of lex_core.cuh into an extra
// token region -- per statement, so one odd statement costs one thread a few hundred bytes of
// sequential work, not the batch its speed.
//
// Everything is NUTDB_HD and window-explicit, so tests/emul runs the identical logic on the host by
// looping over windows and passing the carries along sequentially.
#pragma once
#include "lex_core.cuh"

namespace nlex2 {

using namespace nlex;

enum : uint16_t {
  K_SQ = 1, K_DQ = 2, K_BT = 4, K_NL = 8, K_BS = 16, K_DASH = 32, K_SLASH = 64, K_STAR = 128,
  K_L = 256, K_D = 512, K_DOT = 1024, K_OP = 2048, K_P = 4096, K_WS = 8192,
  K_IE = 16384,  // may follow an identifier (tokenizer/mod.rs:486-503)
  K_NE = 32768   // may follow a numeric literal / query parameter (tokenizer/mod.rs:506-543)
};

struct Lex2Tables {
  uint16_t cls[256];
};

inline void build_lex2_tables(Lex2Tables& T) {
  for (int i = 0; i < 256; i++) {
    uint16_t k = 0;  // '@' '$' '#' '?' '\\' controls, DEL, bytes >= 0x80: no class (invalid in code or left to the exact path)
    uint8_t c = (uint8_t)i;
    if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_') k = K_L;
    else if (c >= '0' && c <= '9') k = K_D;
    else
      switch (c) {
        case '\'': k = K_SQ; break;
        case '"': k = K_DQ; break;
        case '`': k = K_BT; break;
        case '\n': case '\r': k = K_NL | K_WS; break;
        case ' ': case '\t': k = K_WS; break;
        case '\\': k = K_BS; break;
        case '-': k = K_DASH; break;
        case '/': k = K_SLASH; break;
        case '*': k = K_STAR | K_P; break;
        case '.': k = K_DOT; break;
        case '<': case '>': case '=': case '!': k = K_OP; break;
        case '(': case ')': case '[': case ']': case '{': case '}': case ',': case ':': case '+': case '%':
        case '&': case '|': case '^': case '~': case ';': k = K_P; break;
        default: break;
      }
    {
      const char* ident_end = "+-*/%&|^><=!.,;[](){}\t\n\r ";
      const char* num_end = "+-*/%&|^><=!,:;])}\t\n\r ";
      if (c != 0 && std::strchr(ident_end, c)) k |= K_IE;
      if (c != 0 && std::strchr(num_end, c)) k |= K_NE;
    }
    T.cls[i] = k;
  }
}

// raw class masks of one 32-byte window (bit i = byte base+i); bytes at or beyond the batch end are 0
struct Win {
  uint32_t sq, dq, bt, nl, bs, dash, slash, star, L, D, DOT, OP, P, WS, IE, NE;
  uint32_t u;      // bytes 'u' (lex3 only, 0 in the three-pass lexer): an escaped one inside a literal is a backslash-u escape
  uint32_t bnd;    // a statement starts at this byte; also set at position n (virtual end) if inside the window
  uint32_t valid;  // bytes that exist (< n)
};

// what the next window starts with (needed by the last lane / last event of this window)
struct Next {
  uint8_t byte;   // 0 if there is none
  uint8_t bnd;    // a statement starts there (or it is the batch end)
  uint16_t cls;
};

NUTDB_HD uint8_t decay(uint8_t s) { return s == A_CX ? (uint8_t)A_C : (s == A_BC0 ? (uint8_t)A_BC : s); }
NUTDB_HD uint32_t bits_range(int lo, int hi) {  // bits lo..hi inclusive, 0 <= lo, hi <= 31
  if (hi < lo) return 0u;
  const uint32_t up = hi >= 31 ? 0xFFFFFFFFu : ((2u << hi) - 1u);
  return up & ~((1u << lo) - 1u);
}
NUTDB_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
NUTDB_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
NUTDB_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return x ? __builtin_clzll(x) : 64;
#endif
}
NUTDB_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// lane i is preceded by an odd run of backslashes (run may continue from the previous window)
NUTDB_HD bool lane_esc(uint32_t bs, int lane, uint8_t carry_in) {
  if (lane == 0) return carry_in != 0;
  const uint32_t y = bs << (32 - lane);  // bit 31 = byte lane-1
  const int run = clz32(~y);             // consecutive backslashes right below `lane`
  if (run >= lane) return ((lane & 1) != 0) != (carry_in != 0);
  return (run & 1) != 0;
}
NUTDB_HD uint8_t esc_carry_out(uint32_t bs, uint8_t carry_in) {
  const int run = clz32(~bs);  // backslashes ending at bit 31
  if (run >= 32) return carry_in;  // 32 more: parity unchanged
  return (uint8_t)(run & 1);
}

struct Events {
  uint32_t quotes, dd, slst, stsl, all, own;
  uint32_t uesc;  // escaped 'u' bytes
};
NUTDB_HD Events make_events(const Win& w, uint32_t esc, uint8_t prev) {
  Events e;
  e.quotes = (w.sq | w.dq) & ~esc;
  e.dd = w.dash & ((w.dash << 1) | (prev == '-' ? 1u : 0u)) & ~w.bnd;
  e.slst = w.star & ((w.slash << 1) | (prev == '/' ? 1u : 0u)) & ~w.bnd;
  e.stsl = w.slash & ((w.star << 1) | (prev == '*' ? 1u : 0u)) & ~w.bnd;
  e.own = (e.quotes | w.bt | w.nl | e.dd | e.slst | e.stsl) & w.valid;
  e.all = e.own | w.bnd;
  e.uesc = esc & w.u;
  return e;
}
NUTDB_HD uint8_t event_type(const Win& w, const Events& ev, int e) {
  const uint32_t b = 1u << e;
  if (ev.quotes & b) return (w.sq & b) ? (uint8_t)EV_SQ : (uint8_t)EV_DQ;
  if (w.bt & b) return EV_BT;
  if (w.nl & b) return EV_NL;
  if (ev.dd & b) return EV_DD;
  if (ev.slst & b) return EV_SLST;
  return EV_STSL;
}

// ---- phase 1: transition FUNCTION of the context automaton over one window (8 entry states at once) ----
NUTDB_HD uint32_t ctx_window_fn(const LexTables& T, const Win& w, const Events& ev, uint32_t run) {
  uint32_t todo = ev.all;
  int last = -1;
  while (todo) {
    const int e = ctz32(todo);
    todo &= todo - 1;
    if ((w.bnd >> e) & 1u) {
      run = 0u;  // constant function -> A_C: a statement starts here
      last = -100;
      if (!((ev.own >> e) & 1u)) continue;
    }
    if (e > last + 1 && last >= -1) run = vec8_then_row(run, T.a_row[EV_OTHER][0], T.a_row[EV_OTHER][1]);
    const uint8_t t = event_type(w, ev, e);
    run = vec8_then_row(run, T.a_row[t][0], T.a_row[t][1]);
    last = e;
  }
  if (last < 31 && last >= -1) run = vec8_then_row(run, T.a_row[EV_OTHER][0], T.a_row[EV_OTHER][1]);
  return run;
}

// positions preceded by an odd run of backslashes, for all 32 bytes at once
NUTDB_HD uint32_t esc_mask32(uint32_t bs, uint8_t carry_in) {
  uint32_t m = 0;
  // runs are short and rare: walk the run starts
  uint32_t starts = bs & ~(bs << 1);
  if (carry_in && !(bs & 1u)) m |= 1u;  // the run ended with the previous window
  while (starts) {
    const int p = ctz32(starts);
    starts &= starts - 1;
    const uint32_t from = bs >> p;               // run begins at bit 0 of `from`
    int len = ctz32(~from);                      // its length (ctz32(0) cannot happen: ~from has a zero only if from is all ones)
    if (from == 0xFFFFFFFFu) len = 32;
    const int total = len + ((p == 0 && carry_in) ? 1 : 0);  // continued from the previous window
    const int end = p + len;                      // first byte after the run
    if (end < 32 && (total & 1)) m |= 1u << end;
  }
  return m;
}

// ---- stage 2: concrete walk of the context automaton over one window (entry state known) ----
#define NUTDB_L2_NCAP 4
struct StrCarry {  // what later windows need to know about a string / quoted identifier still open
  uint8_t has_open = 0;   // one was opened in this window (and that is the last opening)
  uint8_t esc = 0;        // escaped-flag since that opening (or over the whole window if none)
  uint8_t chk = 0;        // ... a backslash-u escape since that opening: the literal needs the parser's validation
  uint32_t open_pos = 0;  // absolute offset of the opening quote
};
NUTDB_HD StrCarry str_then(const StrCarry& a, const StrCarry& b) {
  if (b.has_open) return b;
  StrCarry r = a;
  r.esc = (uint8_t)(a.esc | b.esc);
  r.chk = (uint8_t)(a.chk | b.chk);
  return r;
}
struct WinCtx {
  uint32_t ct = 0;        // code bytes that may be (part of) a code token
  uint32_t in_str = 0;    // bytes lexed inside '..' / ".."
  uint32_t in_bt = 0;     // bytes lexed inside `..`
  uint32_t close = 0;     // final closing quote of a string / quoted identifier: a token ends here
  uint32_t bad = 0;       // the statement containing this byte needs the exact path
  uint32_t bad_prev = 0;  // the statement ENDING right before this (statement start) byte needs the exact path
  uint32_t escm = 0;
  uint8_t s_out = A_C;
  uint8_t ncap = 0;
  // the literals / quoted identifiers that CLOSE in this window, 16 bits each (two per word):
  //   closing offset[0:5) opening offset[5:10) carried[10] type code[11:13)
  // type code, opened in this window: 0 raw, 1 escaped '..', 2 escaped "..", 3 `..`;
  // opened in an earlier window (carried): the kind 1 = '..', 2 = "..", 3 = `..`
  uint32_t capw[NUTDB_L2_NCAP / 2] = {0, 0};
  uint8_t esc_first = 0;               // this window's escaped-flag contribution before a carried close
  StrCarry sc;
  uint32_t last_bnd1 = 0;              // 1 + absolute offset of the last statement start in the window, 0 if none
};

NUTDB_HD void ctx_window(const Win& w, const Events& ev, uint32_t base, const Next& nx, uint8_t s_in, uint8_t prev_byte,
                         WinCtx& o) {
  uint32_t todo = ev.all, consumed = 0;
  // A newline only matters to a line comment (a_next: every other state maps it to itself or to what it decays to
  // anyway), and a line comment needs a "--" in this window or an open one on entry: otherwise the newline events
  // that are not statement starts are skipped.
  if (s_in != A_LC && ev.dd == 0) todo &= ~(w.nl & ~w.bnd);
  uint8_t s = s_in;
  // '' / "" split exactly at the window start: the closing quote was the previous window's last byte
  const uint8_t b0q = (w.sq & 1u) ? (uint8_t)'\'' : ((w.dq & 1u) ? (uint8_t)'"' : (uint8_t)0);
  int reopen_at = (s_in == A_C && b0q && prev_byte == b0q && !(w.bnd & 1u)) ? 0 : -1;
  int s_pos = -1, p0 = 0;
  uint32_t m_code = 0, m_str = 0, m_bt = 0;
  bool open_local = false;
  uint32_t cur_start = 0;
  uint8_t cur_esc = 0;
  auto assign = [&](int lo, int hi) {
    if (hi < lo) return;
    const uint32_t r = bits_range(lo, hi);
    if (s <= A_CX) m_code |= r;
    else if (s == A_SQ || s == A_DQ) {
      m_str |= r;
      if (w.bs & r) cur_esc = 1;
    } else if (s == A_BT) m_bt |= r;
  };
  auto record = [&](int e, uint8_t local_code, uint8_t kind) {
    o.close |= 1u << e;
    if (o.ncap >= NUTDB_L2_NCAP) {
      o.bad |= 1u << e;
      return;
    }
    const uint32_t c = (uint32_t)e | (((cur_start - base) & 31u) << 5) | (open_local ? 0u : 1u << 10) |
                       ((uint32_t)(open_local ? local_code : kind) << 11);
    o.capw[o.ncap >> 1] |= c << (16u * (o.ncap & 1u));
    if (!open_local) o.esc_first = cur_esc;
    o.ncap++;
  };
  while (todo) {
    const int e = ctz32(todo);
    todo &= todo - 1;
    if ((w.bnd >> e) & 1u) {
      assign(p0, e - 1);
      const uint8_t sa = (e == s_pos + 1) ? s : decay(s);
      if (sa == A_SQ || sa == A_DQ || sa == A_BT || sa == A_BC0 || sa == A_BC) o.bad_prev |= 1u << e;
      s = A_C;
      s_pos = -100;
      reopen_at = -1;
      cur_esc = 0;
      open_local = true;  // nothing can be carried into a new statement
      o.last_bnd1 = base + (uint32_t)e + 1u;
      p0 = e;
      if (!((ev.own >> e) & 1u)) continue;
    }
    assign(p0, e);
    p0 = e + 1;
    const uint8_t a0 = (e == s_pos + 1) ? s : decay(s);
    const uint8_t t = event_type(w, ev, e);
    const uint8_t a1 = a_next(a0, t);
    if (a0 <= A_CX) {
      if (t == EV_SQ || t == EV_DQ) {
        if (reopen_at == e) cur_esc = 1;  // second half of '' / ""
        else {
          open_local = true;
          cur_start = base + (uint32_t)e;
          cur_esc = 0;
          o.sc.has_open = 1;
          o.sc.open_pos = cur_start;
        }
      } else if (t == EV_BT) {
        open_local = true;
        cur_start = base + (uint32_t)e;
        cur_esc = 0;
        o.sc.has_open = 1;
        o.sc.open_pos = cur_start;
      } else if ((t == EV_DD || t == EV_SLST) && a0 == A_C) {
        consumed |= 1u << e;  // second byte of "--" / "/*": not a token
      }
    } else if ((a0 == A_SQ && t == EV_SQ) || (a0 == A_DQ && t == EV_DQ)) {
      const uint32_t same = t == EV_SQ ? w.sq : w.dq;
      bool twin;
      if (e < 31) twin = ((same >> (e + 1)) & 1u) && !((w.bnd >> (e + 1)) & 1u);
      else twin = !nx.bnd && nx.byte == (t == EV_SQ ? '\'' : '"');
      if (twin) {
        reopen_at = e + 1;
      } else {
        record(e, cur_esc ? (t == EV_SQ ? 1 : 2) : 0, t == EV_SQ ? 1 : 2);
        cur_esc = 0;
      }
    } else if (a0 == A_BT && t == EV_BT) {
      record(e, 3, 3);
      // `` : Incomplete (tokenizer/mod.rs:323): the previous byte is the opening backtick
      if (e > 0 ? ((w.bt >> (e - 1)) & 1u) != 0 : prev_byte == '`') o.bad |= 1u << e;
      cur_esc = 0;
    }
    s = a1;
    s_pos = e;
  }
  assign(p0, 31);
  if (s_pos < 31) s = decay(s);
  o.s_out = s;
  o.sc.esc = cur_esc;
  o.ct = m_code & ~consumed & w.valid;
  o.in_str = m_str & w.valid;
  o.in_bt = m_bt & w.valid;
}

// What the emitting pass needs of a window's context walk, packed by the counting pass so that the walk is done once:
//   w0 = ct, w1 = close,
//   w2 = ncap[0:3) has_open[3] sc.esc[4] esc_first[5] open offset[6:11) last statement start offset + 1 [11:17),
//   w3 / w4 = WinCtx::capw.
struct WinCtxPacked {
  uint32_t w[5];
};
NUTDB_HD WinCtxPacked ctx_pack(const WinCtx& o, uint32_t base) {
  WinCtxPacked p;
  p.w[0] = o.ct;
  p.w[1] = o.close;
  p.w[2] = (uint32_t)o.ncap | ((uint32_t)o.sc.has_open << 3) | ((uint32_t)(o.sc.esc != 0) << 4) |
           ((uint32_t)(o.esc_first != 0) << 5) | (((o.sc.open_pos - base) & 31u) << 6) |
           ((o.last_bnd1 ? o.last_bnd1 - base : 0u) << 11);
  p.w[3] = o.capw[0];
  p.w[4] = o.capw[1];
  return p;
}
NUTDB_HD void ctx_unpack(const WinCtxPacked& p, uint32_t base, WinCtx& o) {
  o.ct = p.w[0];
  o.close = p.w[1];
  o.ncap = (uint8_t)(p.w[2] & 7u);
  o.sc.has_open = (uint8_t)((p.w[2] >> 3) & 1u);
  o.sc.esc = (uint8_t)((p.w[2] >> 4) & 1u);
  o.esc_first = (uint8_t)((p.w[2] >> 5) & 1u);
  o.sc.open_pos = base + ((p.w[2] >> 6) & 31u);
  const uint32_t lb = (p.w[2] >> 11) & 63u;
  o.last_bnd1 = lb ? base + lb : 0u;
  o.capw[0] = p.w[3];
  o.capw[1] = p.w[4];
}

// ---- stage 3: what token (if any) ENDS at this lane's byte ----
struct Hist {  // class masks of the previous window, restricted to code-token bytes
  uint32_t L = 0, D = 0, DOT = 0, OP = 0, bnd = 0;
};
struct LaneTok {
  uint8_t has = 0;     // a token ends here
  uint8_t type = 0;
  uint8_t kw = 0;
  uint8_t eof = 0;     // the statement ends after this byte: EOF token follows
  uint8_t bad = 0;     // statement needs the exact path
  uint32_t start = 0;  // absolute
  uint32_t end = 0;
};

// start of the run ending at 64-bit position pos, where cont bit q = "byte q continues byte q-1"; -1 = beyond look-back
NUTDB_HD int run_start(uint64_t cont, int pos) {
  const uint64_t below = pos >= 63 ? ~0ull : ((2ull << pos) - 1ull);
  const uint64_t stop = ~cont & below;  // positions <= pos that do NOT continue
  if (!stop) return -1;
  const int r = 63 - clz64(stop);
  return r == 0 ? -1 : r;  // position 0 always "stops" (nothing is known below it): the run may extend further back
}

// per-window constants of code-token recognition: the thread's own masks on top of the previous window's
struct WinTok {
  uint64_t L64, D64, DOT64, W64, bnd64, contW, contO;
  uint32_t nbnd;     // bit i: byte i+1 starts a statement (or is the batch end)
  uint32_t nextW, nextD, nextDOT, nextOP;  // bit i: byte i+1 has that class and belongs to the same statement
  uint32_t endI, endN;   // bit i: byte i+1 may follow an identifier / a number (or the statement ends there)
  uint32_t EI, EX;       // ends of runs of [A-Za-z0-9_]: identifiers decided by masks alone / numbers and odd cases
};
NUTDB_HD WinTok make_wintok(const Win& w, uint32_t ct, const Hist& h, const Next& nx) {
  WinTok k;
  const uint32_t ctL = w.L & ct, ctD = w.D & ct, ctDOT = w.DOT & ct, ctOP = w.OP & ct;
  k.L64 = ((uint64_t)ctL << 32) | h.L;
  k.D64 = ((uint64_t)ctD << 32) | h.D;
  k.DOT64 = ((uint64_t)ctDOT << 32) | h.DOT;
  const uint64_t OP64 = ((uint64_t)ctOP << 32) | h.OP;
  k.bnd64 = ((uint64_t)w.bnd << 32) | h.bnd;
  k.W64 = k.L64 | k.D64;
  k.contW = k.W64 & (k.W64 << 1) & ~k.bnd64;
  k.contO = OP64 & (OP64 << 1) & ~k.bnd64;
  k.nbnd = (w.bnd >> 1) | ((uint32_t)(nx.bnd != 0) << 31);
  // raw class of the next byte is enough: inside code a word / digit / dot / operator byte cannot change the context
  k.nextW = (((w.L | w.D) >> 1) | ((uint32_t)((nx.cls & (K_L | K_D)) != 0) << 31)) & ~k.nbnd;
  k.nextD = ((w.D >> 1) | ((uint32_t)((nx.cls & K_D) != 0) << 31)) & ~k.nbnd;
  k.nextDOT = ((w.DOT >> 1) | ((uint32_t)((nx.cls & K_DOT) != 0) << 31)) & ~k.nbnd;
  k.nextOP = ((w.OP >> 1) | ((uint32_t)((nx.cls & K_OP) != 0) << 31)) & ~k.nbnd;
  k.endI = ((w.IE >> 1) | ((uint32_t)((nx.cls & K_IE) != 0) << 31)) | k.nbnd;
  k.endN = ((w.NE >> 1) | ((uint32_t)((nx.cls & K_NE) != 0) << 31)) | k.nbnd;
  // Word runs end where the next byte is no word byte.  Runs that start with a letter inside the look-back
  // are identifiers and need nothing but masks; runs that start with a digit (numbers), at the very first
  // look-back position (start unknown) or that a statement start cuts in two go through tok_word().
  const uint32_t E = ct & (w.L | w.D) & ~k.nextW;
  if (k.W64 & (k.W64 << 1) & k.bnd64) {
    k.EX = E;
  } else {
    const uint64_t starts = k.W64 & ~(k.W64 << 1);
    const uint64_t sel = starts & (k.D64 | 1ull);
    const uint64_t sum = k.W64 + sel;  // the carry runs through each selected run and lands right behind it
    uint64_t ends = (sum & ~k.W64) >> 1;
    if (sum < k.W64) ends |= 1ull << 63;  // a selected run ends with the window's last byte
    k.EX = E & (uint32_t)(ends >> 32);
  }
  k.EI = E & ~k.EX;
  return k;
}
template <class Src>
NUTDB_HD uint8_t next_byte(Src& src, const WinTok& k, const Next& nx, uint32_t base, int i) {
  if ((k.nbnd >> i) & 1u) return 0;
  return i < 31 ? src.byte(base + (uint32_t)i + 1u) : nx.byte;
}

// word / number whose LAST byte is byte i of the window (i is the end of a run of [A-Za-z0-9_])
template <class Src>
NUTDB_HD LaneTok tok_word(const LexTables& T, Src& src, const WinTok& k, const Next& nx, uint32_t base, int i) {
  LaneTok r;
  const uint32_t pos = base + (uint32_t)i;
  const int p = 32 + i;
  const int st = run_start(k.contW, p);
  if (st < 0) { r.bad = 1; return r; }
  const uint32_t abs_st = base + (uint32_t)st - 32u;
  if (!((k.D64 >> st) & 1ull)) {  // identifier / keyword (tokenizer/mod.rs:262-282)
    if (!((k.endI >> i) & 1u)) { r.bad = 1; return r; }
    r.has = 1;
    r.type = NUTDB_TT_KeywordOrIdentifier;
    r.start = abs_st;
    r.end = pos + 1;
    return r;
  }
  const uint64_t span = ((p >= 63 ? ~0ull : ((2ull << p) - 1ull))) & ~((1ull << st) - 1ull);
  const bool left_dot = ((k.DOT64 >> (st - 1)) & 1ull) && !((k.bnd64 >> st) & 1ull);
  if (k.L64 & span) {
    // A digit-led run with letters in it is an error (1abc) -- or a hex literal: "0" then x / X then hex digits
    // (tokenizer/mod.rs:201-208; the span is the digits, and there is NO end-of-token check).  Only the plain form
    // is lexed here: every character after "0x" a hex digit and no '.' on either side; "0x1G" (a hex literal
    // followed by an identifier), "0x1.5" and ".0x1" split differently and go to the exact path.
    if (p > st && !left_dot && !((k.nextDOT >> i) & 1u) && src.byte(abs_st) == '0' && (src.byte(abs_st + 1u) | 0x20) == 'x') {
      bool all_hex = true;
      for (uint32_t q = abs_st + 2u; q <= pos; q++) {
        const uint8_t c = src.byte(q);
        all_hex = all_hex && ((c >= '0' && c <= '9') || ((c | 0x20) >= 'a' && (c | 0x20) <= 'f'));
      }
      if (all_hex) {
        r.has = 1;
        r.type = NUTDB_TT_HexLiteral;
        r.start = abs_st + 2u;
        r.end = pos + 1u;
        const uint32_t len = r.end - r.start;
        r.kw = (uint8_t)(len > 255u ? 255u : len);
        return r;
      }
    }
    r.bad = 1;
    return r;
  }
  if ((k.nextDOT >> i) & 1u) {  // digits '.' ...: the token ends later
    if (left_dot) r.bad = 1;    // second dot of one numeric token: error
    return r;
  }
  if (!((k.endN >> i) & 1u)) { r.bad = 1; return r; }
  r.has = 1;
  r.end = pos + 1;
  if (!left_dot) {  // integer literal (tokenizer/mod.rs:196-238)
    r.type = NUTDB_TT_IntegerLiteral;
    r.start = abs_st;
    const uint32_t len = r.end - r.start;
    r.kw = (uint8_t)(len > 255u ? 255u : len);
    return r;
  }
  // float: [digits] '.' digits (tokenizer/mod.rs:246-258)
  const int d = st - 1;
  int fs = d;
  if (d > 0 && ((k.W64 >> (d - 1)) & 1ull) && !((k.bnd64 >> d) & 1ull)) {
    const int ls = run_start(k.contW, d - 1);
    if (ls < 0) { r.bad = 1; return r; }
    if ((k.D64 >> ls) & 1ull) {  // digits before the dot belong to the literal
      fs = ls;
      if (((k.DOT64 >> (ls - 1)) & 1ull) && !((k.bnd64 >> ls) & 1ull)) r.bad = 1;  // 1.2.3
    }
  } else if (d == 0) {
    r.bad = 1;  // cannot see what precedes the dot
  }
  r.type = NUTDB_TT_FloatLiteral;
  r.start = base + (uint32_t)fs - 32u;
  return r;
}

// byte i is a '.' in code
template <class Src>
NUTDB_HD LaneTok tok_dot(const LexTables& T, Src& src, const WinTok& k, const Next& nx, uint32_t base, int i) {
  LaneTok r;
  const uint32_t pos = base + (uint32_t)i;
  const int p = 32 + i;
  if ((k.nextD >> i) & 1u) return r;  // '.' digits: ends at the last digit
  bool is_float = false;
  int fs = p;
  if (((k.W64 >> (p - 1)) & 1ull) && !((k.bnd64 >> p) & 1ull)) {
    const int ls = run_start(k.contW, p - 1);
    if (ls < 0) { r.bad = 1; return r; }
    if ((k.D64 >> ls) & 1ull) {  // digits '.'  (a word with letters in it before the dot is flagged at its own end)
      is_float = true;
      fs = ls;
      if (((k.DOT64 >> (ls - 1)) & 1ull) && !((k.bnd64 >> ls) & 1ull)) r.bad = 1;  // .5.
    }
  }
  r.has = 1;
  r.end = pos + 1;
  if (is_float) {
    if (!((k.endN >> i) & 1u)) { r.bad = 1; return r; }
    r.type = NUTDB_TT_FloatLiteral;
    r.start = base + (uint32_t)fs - 32u;
  } else {
    r.type = NUTDB_TT_Dot;  // no end check (tokenizer/mod.rs:248-250)
    r.start = pos;
  }
  return r;
}

// byte i is one of < > = ! in code: at most two in a row are handled here (tokenizer/mod.rs:393-428)
template <class Src>
NUTDB_HD LaneTok tok_op(Src& src, const WinTok& k, const Next& nx, uint32_t base, int i, uint8_t prev_byte) {
  LaneTok r;
  const uint32_t pos = base + (uint32_t)i;
  const int p = 32 + i;
  const uint8_t b = src.byte(pos);
  const int st = run_start(k.contO, p);
  if (st < 0 || p - st >= 2) { r.bad = 1; return r; }
  const bool more = ((k.nextOP >> i) & 1u) != 0;
  auto pair_type = [](uint8_t c0, uint8_t c1) -> uint8_t {
    if (c0 == '<') return c1 == '=' ? NUTDB_TT_LtEq : (c1 == '>' ? NUTDB_TT_NotEq : (c1 == '<' ? NUTDB_TT_BitLShift : 0));
    if (c0 == '>') return c1 == '=' ? NUTDB_TT_GtEq : (c1 == '>' ? NUTDB_TT_BitRShift : 0);
    if (c0 == '!') return c1 == '=' ? NUTDB_TT_NotEq : 0;
    return 0;
  };
  auto single_type = [](uint8_t c0) -> uint8_t {
    return c0 == '<' ? NUTDB_TT_Lt : (c0 == '>' ? NUTDB_TT_Gt : (c0 == '=' ? NUTDB_TT_Eq : 0xFF));
  };
  if (p - st == 0) {
    if (more && pair_type(b, next_byte(src, k, nx, base, i))) return r;  // first half of a two-character operator
    const uint8_t ty = single_type(b);
    if (ty == 0xFF) { r.bad = 1; return r; }  // lone '!'
    r.has = 1;
    r.type = ty;
    r.start = pos;
    r.end = pos + 1;
    return r;
  }
  if (more) { r.bad = 1; return r; }  // three or more
  const uint8_t pb = i > 0 ? src.byte(pos - 1) : prev_byte;
  const uint8_t pt = pair_type(pb, b);
  r.has = 1;
  r.end = pos + 1;
  if (pt) {
    r.type = pt;
    r.start = pos - 1;
  } else {
    const uint8_t ty = single_type(b);
    if (ty == 0xFF) { r.bad = 1; r.has = 0; return r; }
    r.type = ty;
    r.start = pos;
  }
  return r;
}

// statements that need the exact path, as far as masks can tell (token-level checks add to this)
template <class Src, class Ctx>
NUTDB_HD uint32_t win_bad_mask(Src& src, const Win& w, const Ctx& o, uint32_t base, uint8_t prev_byte) {
  const uint32_t known = w.sq | w.dq | w.bt | w.WS | w.dash | w.slash | w.L | w.D | w.DOT | w.OP | w.P;
  uint32_t bad = o.bad;
  bad |= o.ct & w.valid & ~known;              // '@' '$' '#' '?' '\\' controls, non-ASCII in code
  bad |= o.ct & (w.sq | w.dq) & o.escm;        // an escaped quote outside a string
  bad |= o.in_bt & w.nl;                       // raw newline in `..` (tokenizer/mod.rs:336)
  uint32_t nls = o.in_str & w.nl;              // raw CR / LF inside a string literal must be escaped (mod.rs:147-170)
  while (nls) {
    const int i = ctz32(nls);
    nls &= nls - 1;
    const uint32_t bit = 1u << i;
    if (o.escm & bit) continue;
    const uint8_t b = src.byte(base + (uint32_t)i);
    if (b == '\n') {
      const uint8_t pb = i > 0 ? src.byte(base + (uint32_t)i - 1u) : prev_byte;
      const bool pesc = i > 0 ? ((o.escm >> (i - 1)) & 1u) != 0 : false;
      if (pb == '\r' && pesc) continue;  // backslash CR LF
    }
    bad |= bit;
  }
  return bad;
}

// Tokens are extracted CLASS BY CLASS (all single-character tokens, then all words/numbers, ...), not in
// position order: every loop below runs the same code in all lanes of a warp.  The index of a token only
// needs the mask of ALL positions where a token ends ("has"), which the counting pass computes and stores.

// single-character punctuation and lone '-' '/' need no look at anything but masks
NUTDB_HD uint32_t simple_token_mask(const Win& w, const WinCtx& o, const Next& nx, const WinTok& k) {
  const uint32_t nextDash = ((w.dash >> 1) | ((uint32_t)(nx.byte == '-') << 31)) & ~k.nbnd;
  const uint32_t nextStar = ((w.star >> 1) | ((uint32_t)(nx.byte == '*') << 31)) & ~k.nbnd;
  return o.ct & (w.P | (w.dash & ~nextDash) | (w.slash & ~nextStar));
}

// counting pass: positions where a token ENDS (code tokens and literal closes; EOF tokens are counted from
// the boundary mask) and positions whose statement needs the exact path
template <class Src>
NUTDB_HD uint32_t win_has_mask(const LexTables& T, Src& src, const Win& w, const WinCtx& o, const Hist& h, const Next& nx,
                               uint32_t base, uint8_t prev_byte, uint32_t& bad) {
  bad = win_bad_mask(src, w, o, base, prev_byte);
  const WinTok k = make_wintok(w, o.ct, h, nx);
  uint32_t has = simple_token_mask(w, o, nx, k) | o.close;
  has |= k.EI & k.endI;    // identifiers / keywords: the byte after must be allowed to follow one
  bad |= k.EI & ~k.endI;
  uint32_t todo = k.EX;    // numbers, runs with unknown start
  while (todo) {
    const int i = ctz32(todo);
    todo &= todo - 1;
    const LaneTok t = tok_word(T, src, k, nx, base, i);
    has |= (uint32_t)t.has << i;
    bad |= (uint32_t)t.bad << i;
  }
  todo = o.ct & w.DOT;
  while (todo) {
    const int i = ctz32(todo);
    todo &= todo - 1;
    const LaneTok t = tok_dot(T, src, k, nx, base, i);
    has |= (uint32_t)t.has << i;
    bad |= (uint32_t)t.bad << i;
  }
  todo = o.ct & w.OP;
  while (todo) {
    const int i = ctz32(todo);
    todo &= todo - 1;
    const LaneTok t = tok_op(src, k, nx, base, i, prev_byte);
    has |= (uint32_t)t.has << i;
    bad |= (uint32_t)t.bad << i;
  }
  return has;
}
NUTDB_HD uint32_t win_eof_mask(const Win& w, const Next& nx) {
  return ((w.bnd >> 1) | ((uint32_t)(nx.bnd != 0) << 31)) & w.valid;
}

// Sink: void token(uint32_t index, uint8_t type, uint32_t start_rel, uint32_t end_rel, uint8_t kw);
//       void stmt_begin(uint32_t abs_pos, uint32_t first_index); void stmt_end(uint32_t abs_last_byte, uint32_t end_index);
// emitting pass: `has` is the mask stored by the counting pass, `index` the index of the window's first token
template <class Src, class Sink>
NUTDB_HD void win_emit(const LexTables& T, Src& src, Sink& sink, const Win& w, const WinCtx& o, const Hist& h, const Next& nx,
                       uint32_t base, uint8_t prev_byte, const StrCarry& sc_in, uint32_t stmt_start_in, uint32_t index,
                       uint32_t has) {
  const uint32_t eofm = win_eof_mask(w, nx);
  const WinTok k = make_wintok(w, o.ct, h, nx);
  const uint32_t bnds = w.bnd & w.valid;
  auto index_of = [&](int i) {  // tokens (and EOF tokens) that end before byte i
    const uint32_t below = (1u << i) - 1u;
    return index + (uint32_t)popc32(has & below) + (uint32_t)popc32(eofm & below);
  };
  auto stmt_of = [&](int i) {  // start of the statement byte i belongs to
    const uint32_t bb = bnds & (i >= 31 ? 0xFFFFFFFFu : ((2u << i) - 1u));
    return bb ? base + (uint32_t)(31 - clz32(bb)) : stmt_start_in;
  };
  // statement starts and ends
  uint32_t todo = bnds;
  while (todo) {
    const int i = ctz32(todo);
    todo &= todo - 1;
    sink.stmt_begin(base + (uint32_t)i, index_of(i));
  }
  todo = eofm;
  while (todo) {
    const int i = ctz32(todo);
    todo &= todo - 1;
    const uint32_t sst = stmt_of(i), e = index_of(i) + ((has >> i) & 1u);
    sink.token(e, NUTDB_TT_EOF, base + (uint32_t)i + 1u - sst, base + (uint32_t)i + 1u - sst, 0);
    sink.stmt_end(base + (uint32_t)i, e + 1u);
  }
  // single-character tokens
  todo = simple_token_mask(w, o, nx, k);
  while (todo) {
    const int i = ctz32(todo);
    todo &= todo - 1;
    const uint32_t pos = base + (uint32_t)i, sst = stmt_of(i);
    const uint8_t b = src.byte(pos);
    const uint8_t type = b == '-' ? (uint8_t)NUTDB_TT_Minus : (b == '/' ? (uint8_t)NUTDB_TT_Div : T.single_tt[b]);
    sink.token(index_of(i), type, pos - sst, pos + 1u - sst, 0);
  }
  // identifiers / keywords: the start is the highest run start at or below the end
  {
    const uint64_t starts = k.W64 & ~(k.W64 << 1);
    todo = k.EI & has;
    while (todo) {
      const int i = ctz32(todo);
      todo &= todo - 1;
      const int p = 32 + i;
      const int st = 63 - clz64(starts & (p >= 63 ? ~0ull : ((2ull << p) - 1ull)));
      const uint32_t start = base + (uint32_t)st - 32u, end = base + (uint32_t)i + 1u, sst = stmt_of(i);
      const uint32_t len = end - start;
      uint8_t kw = 0;
      if (len >= 2 && len <= 10) {
        const uint8_t* wp = src.span(start, len);  // the word as contiguous bytes (almost always)
        if (wp) {
          uint32_t w0, w1, w2;
          load_word12(wp, len, w0, w1, w2);
          kw = keyword_lookup_words(T, len, w0, w1, w2);
        } else {
          Src& sr = src;
          kw = keyword_lookup(T, len, [&sr, start](uint32_t q) { return sr.byte(start + q); });
        }
      }
      sink.token(index_of(i), NUTDB_TT_KeywordOrIdentifier, start - sst, end - sst, kw);
    }
  }
  // numbers (and the odd word runs)
  todo = k.EX & has;
  while (todo) {
    const int i = ctz32(todo);
    todo &= todo - 1;
    const LaneTok t = tok_word(T, src, k, nx, base, i);
    const uint32_t sst = stmt_of(i);
    uint8_t kw = t.kw;
    if (t.type == NUTDB_TT_KeywordOrIdentifier) {
      const uint32_t len = t.end - t.start;
      if (len >= 2 && len <= 10) {
        Src& sr = src;
        const uint32_t s0 = t.start;
        kw = keyword_lookup(T, len, [&sr, s0](uint32_t q) { return sr.byte(s0 + q); });
      }
    }
    sink.token(index_of(i), t.type, t.start - sst, t.end - sst, kw);
  }
  todo = o.ct & w.DOT & has;
  while (todo) {
    const int i = ctz32(todo);
    todo &= todo - 1;
    const LaneTok t = tok_dot(T, src, k, nx, base, i);
    const uint32_t sst = stmt_of(i);
    sink.token(index_of(i), t.type, t.start - sst, t.end - sst, 0);
  }
  todo = o.ct & w.OP & has;
  while (todo) {
    const int i = ctz32(todo);
    todo &= todo - 1;
    const LaneTok t = tok_op(src, k, nx, base, i, prev_byte);
    const uint32_t sst = stmt_of(i);
    sink.token(index_of(i), t.type, t.start - sst, t.end - sst, 0);
  }
  // strings and quoted identifiers, at their closing quote
  for (uint32_t c = 0; c < o.ncap; c++) {
    const uint32_t cw = (o.capw[c >> 1] >> (16u * (c & 1u))) & 0xFFFFu;
    const int i = (int)(cw & 31u);
    const uint32_t pos = base + (uint32_t)i, sst = stmt_of(i);
    const uint32_t code = (cw >> 11) & 3u;
    uint8_t type = code == 0 ? (uint8_t)NUTDB_TT_RawStringLiteral
                   : code == 1 ? (uint8_t)NUTDB_TT_EscapedSQStringLiteral
                   : code == 2 ? (uint8_t)NUTDB_TT_EscapedDQStringLiteral : (uint8_t)NUTDB_TT_DelimitedIdentifier;
    uint32_t start = base + ((cw >> 5) & 31u) + 1u;
    if ((cw >> 10) & 1u) {  // opened in an earlier window: offset and escaped flag come from the carry
      start = sc_in.open_pos + 1u;
      const bool escd = (sc_in.esc | o.esc_first) != 0;
      type = code == 3 ? (uint8_t)NUTDB_TT_DelimitedIdentifier
                       : (escd ? (code == 1 ? (uint8_t)NUTDB_TT_EscapedSQStringLiteral : (uint8_t)NUTDB_TT_EscapedDQStringLiteral)
                               : (uint8_t)NUTDB_TT_RawStringLiteral);
    }
    sink.token(index_of(i), type, start - sst, pos - sst, 0);
  }
  // Closes beyond the capture array: their statement is flagged for the exact path, which gives it a token range of
  // its own; the slots counted for it here are referenced by nobody, but every output byte has to be a function of
  // the input alone (reproducible arrays, nutdb_gpu_batch_hash), so they get a fixed filler.
  if (o.ncap >= NUTDB_L2_NCAP) {
    uint32_t rest = o.close & has;
    for (uint32_t c = 0; c < o.ncap; c++) rest &= ~(1u << ((o.capw[c >> 1] >> (16u * (c & 1u))) & 31u));
    while (rest) {
      const int i = ctz32(rest);
      rest &= rest - 1;
      sink.token(index_of(i), NUTDB_TT_POISON, 0u, 0u, 0);
    }
  }
}

}  // namespace nlex2
