// lex3_core.cuh -- window logic of the SINGLE-PASS lexer (kernel: lex3_kernels.cuh, k_lex3).
//
// Replaces Tokenizer::next_token and its 14 routines (reference src/parser/tokenizer/mod.rs:66-468) for
// valid text; anything the reference rejects is handed, per statement, to the exact walker of lex_core.cuh.
//
// One THREAD owns one 32-byte window.  What changed against lex2 (three passes over the text):
//
//   * class masks by BIT-SLICING: the window's 32 bytes are transposed into 8 bit planes (16 byte
//     permutes + an 8x8 bit-matrix transposition on all four byte lanes at once), after which every
//     character class is a boolean function of the planes evaluated for 32 bytes per instruction
//     (no table look-ups, no shared-memory bank conflicts);
//   * token ENDS and token STARTS are pure mask arithmetic (win_tokens3): identifiers, numbers, floats,
//     dots, one- and two-character operators; the only loops left in the window thread are the sparse
//     context events (strings / comments, lex2_core.cuh) and the enumeration of the token records;
//   * the window thread only writes a RECORD (start, end, flags) per token; type, keyword id, the
//     reference's end-of-token checks (tokenizer/mod.rs:486-543) and the statement-relative spans are
//     computed by token_finish3 with one thread per TOKEN and stored coalesced.
//
// Everything is NUTDB_HD: tests/emul/emul_lex3.cpp runs the identical functions on the host.
#pragma once
#include "lex2_core.cuh"

#if defined(__CUDACC__)
#define NUTDB_UNROLL _Pragma("unroll")
#else
#define NUTDB_UNROLL
#endif

namespace nlex3 {

using namespace nlex;
using namespace nlex2;

// ---- bit-slicing -------------------------------------------------------------------------------------------
NUTDB_HD uint32_t bperm(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, b, sel);
#else
  const uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8u * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
  return r;
#endif
}
NUTDB_HD void swapmove(uint32_t& a, uint32_t& b, uint32_t m, int n) {
  const uint32_t t = ((a >> n) ^ b) & m;
  b ^= t;
  a ^= t << n;
}
// v[0..7] = the window's 32 bytes as little-endian words.  Returns the 8 bit planes: bit i of p[k] = bit k of byte i.
NUTDB_HD void bit_planes(const uint32_t v[8], uint32_t p[8]) {
  // x[j] byte c = byte 8c + j of the window (a 4x4 byte transposition of {v0,v2,v4,v6} and of {v1,v3,v5,v7})
  uint32_t x[8];
NUTDB_UNROLL
  for (int g = 0; g < 2; g++) {
    const uint32_t a = v[g], b = v[2 + g], c = v[4 + g], d = v[6 + g];
    const uint32_t t0 = bperm(a, b, 0x5140), t1 = bperm(a, b, 0x7362);
    const uint32_t u0 = bperm(c, d, 0x5140), u1 = bperm(c, d, 0x7362);
    x[4 * g + 0] = bperm(t0, u0, 0x5410);
    x[4 * g + 1] = bperm(t0, u0, 0x7632);
    x[4 * g + 2] = bperm(t1, u1, 0x5410);
    x[4 * g + 3] = bperm(t1, u1, 0x7632);
  }
  // 8x8 bit-matrix transposition (rows x[j], columns = bit in byte), all four byte lanes at once
NUTDB_UNROLL
  for (int j = 0; j < 4; j++) swapmove(x[j], x[j + 4], 0x0F0F0F0Fu, 4);
NUTDB_UNROLL
  for (int j = 0; j < 8; j += 4) {
    swapmove(x[j], x[j + 2], 0x33333333u, 2);
    swapmove(x[j + 1], x[j + 3], 0x33333333u, 2);
  }
NUTDB_UNROLL
  for (int j = 0; j < 8; j += 2) swapmove(x[j], x[j + 1], 0x55555555u, 1);
NUTDB_UNROLL
  for (int k = 0; k < 8; k++) p[k] = x[k];
}

// the four comparison characters separately (pairs like "<=" are decided by masks)
struct Ops {
  uint32_t LT, GT, EQ, BANG;
};

// class masks of the window from its bit planes (bytes beyond `valid` are cleared)
NUTDB_HD void classify_planes(const uint32_t p[8], uint32_t valid, Win& w, Ops& op) {
  const uint32_t b0 = p[0], b1 = p[1], b2 = p[2], b3 = p[3], b4 = p[4], b5 = p[5], b6 = p[6], n7 = ~p[7] & valid;
  // decoders: (b3, b2) with "ASCII and valid" folded in, (b1, b0), and the eight rows (b6, b5, b4)
  const uint32_t h00 = n7 & ~b3 & ~b2, h01 = n7 & ~b3 & b2, h10 = n7 & b3 & ~b2, h11 = n7 & b3 & b2;
  const uint32_t l00 = ~b1 & ~b0, l01 = ~b1 & b0, l10 = b1 & ~b0, l11 = b1 & b0;
  const uint32_t r0 = ~b6 & ~b5 & ~b4, r2 = ~b6 & b5 & ~b4, r3 = ~b6 & b5 & b4, r4 = b6 & ~b5 & ~b4, r5 = b6 & ~b5 & b4,
                 r6 = b6 & b5 & ~b4, r7 = b6 & b5 & b4;
  const uint32_t tab = r0 & h10 & l01, lf = r0 & h10 & l10, cr = r0 & h11 & l01, sp = r2 & h00 & l00;
  w.nl = lf | cr;
  w.WS = w.nl | tab | sp;
  op.BANG = r2 & h00 & l01;
  w.dq = r2 & h00 & l10;
  w.sq = r2 & h01 & l11;
  w.star = r2 & h10 & l10;
  w.dash = r2 & h11 & l01;
  w.DOT = r2 & h11 & l10;
  w.slash = r2 & h11 & l11;
  op.LT = r3 & h11 & l00;
  op.EQ = r3 & h11 & l01;
  op.GT = r3 & h11 & l10;
  w.bs = r5 & h11 & l00;
  w.u = r7 & h01 & l01;
  w.bt = r6 & h00 & l00;
  w.D = r3 & n7 & (~b3 | (~b2 & ~b1));
  const uint32_t lo_nz = b3 | b2 | b1 | b0, lo_leA = ~b3 | (~b2 & ~l11);
  w.L = (n7 & (((r4 | r6) & lo_nz) | ((r5 | r7) & lo_leA))) | (r5 & h11 & l11);
  // single-character tokens: % & ( ) * + ,   : ;   [ ] ^   { | } ~
  w.P = (r2 & ((h01 & (b1 ^ b0)) | h10 | (h11 & l00))) | (r3 & h10 & b1) | (r5 & ((h10 & l11) | (h11 & (b1 ^ b0)))) |
        (r7 & ((h10 & l11) | (h11 & ~l11)));
  w.OP = op.LT | op.GT | op.EQ | op.BANG;
  // what may follow an identifier (tokenizer/mod.rs:486-503) / a numeric literal (:506-543)
  const uint32_t colon = r3 & h10 & l10, tilde = r7 & h11 & l10, lparen = r2 & h10 & l00, lbracket = r5 & h10 & l11,
                 lbrace = r7 & h10 & l11;
  const uint32_t common = w.WS | w.OP | w.dash | w.slash;
  w.IE = common | w.DOT | (w.P & ~(colon | tilde));
  w.NE = common | (w.P & ~(lparen | lbracket | lbrace | tilde));
  w.valid = valid;
}

// ---- concrete walk of the context automaton over one window (entry state known) ------------------------------
// lex2's ctx_window with the literals that close in the window described by MASKS instead of a capture array, so
// any number of them may close in one window: `openm` = opening quotes seen in this window, `close` = final closing
// quotes, `escd` = closes of escaped literals opened here.  A close without an opening quote below it belongs to the
// literal that was open on entry (at most one per window): its offset and escaped flag come from the carry.
struct WinCtx3 {
  uint32_t ct = 0;        // code bytes that may be (part of) a code token
  uint32_t in_str = 0;    // bytes lexed inside '..' / ".."
  uint32_t in_bt = 0;     // bytes lexed inside `..`
  uint32_t close = 0, openm = 0, escd = 0;
  uint32_t ustr = 0;      // escaped 'u' bytes inside '..' / "..": a literal that holds one has a backslash-u escape
  uint32_t csq = 0, cdq = 0;  // closes by quote character (the rest of `close` are backticks)
  uint32_t bad = 0;       // the statement containing this byte needs the exact path
  uint32_t bad_prev = 0;  // the statement ENDING right before this (statement start) byte needs the exact path
  uint32_t escm = 0;
  uint8_t s_out = A_C;
  uint8_t esc_first = 0;  // this window's escaped-flag contribution before a carried close
  StrCarry sc;
  uint32_t last_bnd1 = 0;  // 1 + absolute offset of the last statement start in the window, 0 if none
};

NUTDB_HD void ctx_window3(const Win& w, const Events& ev, uint32_t base, const Next& nx, uint8_t s_in, uint8_t prev_byte,
                          WinCtx3& o) {
  uint32_t todo = ev.all, consumed = 0;
  // A newline only matters to a line comment, and a line comment needs a "--" in this window or an open one on entry
  if (s_in != A_LC && ev.dd == 0) todo &= ~(w.nl & ~w.bnd);
  uint8_t s = s_in;
  // '' / "" split exactly at the window start: the closing quote was the previous window's last byte
  const uint8_t b0q = (w.sq & 1u) ? (uint8_t)'\'' : ((w.dq & 1u) ? (uint8_t)'"' : (uint8_t)0);
  int reopen_at = (s_in == A_C && b0q && prev_byte == b0q && !(w.bnd & 1u)) ? 0 : -1;
  int s_pos = -1, p0 = 0;
  uint32_t m_code = 0, m_str = 0, m_bt = 0;
  bool open_local = false;
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
  while (todo) {
    const int e = ctz32(todo);
    todo &= todo - 1;
    const uint32_t bit = 1u << e;
    if (w.bnd & bit) {
      assign(p0, e - 1);
      const uint8_t sa = (e == s_pos + 1) ? s : decay(s);
      if (sa == A_SQ || sa == A_DQ || sa == A_BT || sa == A_BC0 || sa == A_BC) o.bad_prev |= bit;
      s = A_C;
      s_pos = -100;
      reopen_at = -1;
      cur_esc = 0;
      open_local = true;  // nothing can be carried into a new statement
      o.last_bnd1 = base + (uint32_t)e + 1u;
      p0 = e;
      if (!(ev.own & bit)) continue;
    }
    assign(p0, e);
    p0 = e + 1;
    const uint8_t a0 = (e == s_pos + 1) ? s : decay(s);
    const uint8_t t = event_type(w, ev, e);
    const uint8_t a1 = a_next(a0, t);
    if (a0 <= A_CX) {
      if (t == EV_SQ || t == EV_DQ || t == EV_BT) {
        if (t != EV_BT && reopen_at == e) {
          cur_esc = 1;  // second half of '' / ""
        } else {
          open_local = true;
          cur_esc = 0;
          o.openm |= bit;
          o.sc.has_open = 1;
          o.sc.open_pos = base + (uint32_t)e;
        }
      } else if ((t == EV_DD || t == EV_SLST) && a0 == A_C) {
        consumed |= bit;  // second byte of "--" / "/*": not a token
      }
    } else if ((a0 == A_SQ && t == EV_SQ) || (a0 == A_DQ && t == EV_DQ)) {
      const uint32_t same = t == EV_SQ ? w.sq : w.dq;
      bool twin;
      if (e < 31) twin = ((same >> (e + 1)) & 1u) && !((w.bnd >> (e + 1)) & 1u);
      else twin = !nx.bnd && nx.byte == (t == EV_SQ ? '\'' : '"');
      if (twin) {
        reopen_at = e + 1;
      } else {
        o.close |= bit;
        if (t == EV_SQ) o.csq |= bit;
        else o.cdq |= bit;
        if (cur_esc) {
          if (open_local) o.escd |= bit;
          else o.esc_first = 1;
        }
        cur_esc = 0;
      }
    } else if (a0 == A_BT && t == EV_BT) {
      o.close |= bit;
      // `` : Incomplete (tokenizer/mod.rs:323): the previous byte is the opening backtick
      if (e > 0 ? ((w.bt >> (e - 1)) & 1u) != 0 : prev_byte == '`') o.bad |= bit;
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
  // backslash-u escapes, by mask: of the literal still open at the end (everything behind its opening quote, or the
  // whole window if it was open on entry; what a literal that closed here leaves in the carry is never used)
  o.ustr = ev.uesc & o.in_str;
  o.sc.chk = (o.ustr & (o.sc.has_open ? ~((2u << (o.sc.open_pos - base)) - 1u) : 0xFFFFFFFFu)) != 0u;
  o.in_bt = m_bt & w.valid;
}
// the first literal that closes in the window was opened in an earlier one
NUTDB_HD bool has_carried_close(const WinCtx3& o) {
  return o.close != 0u && (o.openm & ((o.close & (0u - o.close)) - 1u)) == 0u;
}

// ---- code tokens by mask arithmetic ------------------------------------------------------------------------
struct Hist3 {  // the previous window: code-token class masks (restricted to code bytes) and statement starts
  uint32_t L = 0, D = 0, DOT = 0, bnd = 0;
};
struct TokMasks {
  uint32_t has = 0;    // a token's last byte (code tokens; closing quotes of literals / quoted identifiers)
  uint32_t eofm = 0;   // the statement ends after this byte: an EOF token follows
  uint32_t bad = 0;    // the statement containing this byte needs the exact lexer
  uint32_t bad_prev = 0;  // ... and so does the statement that ends right before this byte
  uint64_t TS = 0;     // first bytes of code tokens over [previous window | this window]
  uint64_t L64 = 0, DOT64 = 0, N64 = 0;  // letters, dots, digits | dots (restricted to code bytes), same 64 positions
};

NUTDB_HD int ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}
NUTDB_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}

// is (c0, c1) one two-character operator (tokenizer/mod.rs:393-428)?
NUTDB_HD uint8_t pair_type(uint8_t c0, uint8_t c1) {
  if (c0 == '<') return c1 == '=' ? (uint8_t)NUTDB_TT_LtEq : (c1 == '>' ? (uint8_t)NUTDB_TT_NotEq : (c1 == '<' ? (uint8_t)NUTDB_TT_BitLShift : (uint8_t)0));
  if (c0 == '>') return c1 == '=' ? (uint8_t)NUTDB_TT_GtEq : (c1 == '>' ? (uint8_t)NUTDB_TT_BitRShift : (uint8_t)0);
  if (c0 == '!') return c1 == '=' ? (uint8_t)NUTDB_TT_NotEq : (uint8_t)0;
  return 0;
}

// prev / prev2: the two bytes in front of the window (0 if none); they only matter for operator pairs, whose bytes
// cannot be the tail of a literal or comment, so their raw values are enough.
// nx.cls: bit 0 = the next window's first byte may follow an identifier, bit 1 = ... a numeric literal.
NUTDB_HD void win_tokens3(const Win& w, const Ops& op, const WinCtx3& o, const Hist3& h, const Next& nx, uint8_t prev,
                          uint8_t prev2, TokMasks& m) {
  const uint32_t ct = o.ct;
  const uint32_t cL = w.L & ct, cD = w.D & ct, cDOT = w.DOT & ct;
  const uint64_t L64 = ((uint64_t)cL << 32) | h.L, D64 = ((uint64_t)cD << 32) | h.D, DOT64 = ((uint64_t)cDOT << 32) | h.DOT;
  const uint64_t bnd64 = ((uint64_t)w.bnd << 32) | h.bnd;
  const uint64_t W64 = L64 | D64;
  const uint64_t adj = W64 & (W64 << 1);
  const uint64_t cont = adj & ~bnd64;  // this byte continues the word run of the byte before it
  const uint64_t Wstart = W64 & ~cont;
  const uint64_t split = adj & bnd64;  // a statement starts in the middle of a run of word characters
  // digit-led runs (numbers).  Adding the run's first bit to the run makes the carry sweep exactly that run.
  uint64_t dled;
  if (split == 0) {
    const uint64_t sum = W64 + (Wstart & D64);
    dled = W64 & ~sum;
  } else {  // two runs touch without a gap (back-to-back statements): run by run
    dled = 0;
    uint64_t st = Wstart & D64;
    while (st) {
      const int s = ctz64(st);
      st &= st - 1;
      const uint64_t follow = s >= 63 ? 0ull : (cont >> (s + 1));
      const int len = 1 + (~follow ? ctz64(~follow) : 63 - s);
      dled |= (len >= 64 ? ~0ull : ((1ull << len) - 1ull)) << s;
    }
  }
  const uint64_t Wend = W64 & ~(cont >> 1);
  const uint64_t joinL = DOT64 & ((dled & Wend) << 1) & ~bnd64;    // "12." : the dot belongs to the number
  const uint64_t frac = Wstart & D64 & (DOT64 << 1) & ~bnd64;      // ".5" / "1.5": the digits belong to the dot
  const uint64_t TSw = (Wstart & ~frac) | (DOT64 & ~joinL);
  // what follows each byte of this window (a statement start cuts everything)
  const uint32_t nbnd = (w.bnd >> 1) | ((uint32_t)(nx.bnd != 0) << 31);
  const uint32_t nxb = nx.bnd ? 0u : (uint32_t)nx.byte;
  const uint32_t nxW = ((nxb >= '0' && nxb <= '9') || ((nxb | 0x20u) >= 'a' && (nxb | 0x20u) <= 'z') || nxb == '_') ? 1u : 0u;
  const uint32_t nxD = (nxb >= '0' && nxb <= '9') ? 1u : 0u;
  const uint32_t nextW = (((w.L | w.D) >> 1) | (nxW << 31)) & ~nbnd;
  const uint32_t nextD = ((w.D >> 1) | (nxD << 31)) & ~nbnd;
  const uint32_t nextDOT = ((w.DOT >> 1) | ((uint32_t)(nxb == '.') << 31)) & ~nbnd;
  const uint32_t nextDash = ((w.dash >> 1) | ((uint32_t)(nxb == '-') << 31)) & ~nbnd;
  const uint32_t nextStar = ((w.star >> 1) | ((uint32_t)(nxb == '*') << 31)) & ~nbnd;
  // words and numbers end where the run ends -- unless a number goes on with '.'; a dot ends unless digits follow
  const uint32_t E = (cL | cD) & ~nextW;
  const uint32_t has_word = E & ~((uint32_t)(dled >> 32) & nextDOT);
  const uint32_t has_dot = cDOT & ~nextD;
  // single-character tokens and lone '-' '/'
  const uint32_t simple = ct & (w.P | (w.dash & ~nextDash) | (w.slash & ~nextStar));
  // < > = ! : second bytes of "<=" ">=" "!=" "<>" "<<" ">>"
  const uint32_t pLT = (op.LT << 1) | (uint32_t)(prev == '<'), pGT = (op.GT << 1) | (uint32_t)(prev == '>'),
                 pBANG = (op.BANG << 1) | (uint32_t)(prev == '!');
  const uint32_t cLT = op.LT & ct, cGT = op.GT & ct, cEQ = op.EQ & ct, cBANG = op.BANG & ct;
  const uint32_t cOP = cLT | cGT | cEQ | cBANG;
  const uint32_t pair2 = ((cEQ & (pLT | pGT | pBANG)) | (cGT & (pLT | pGT)) | (cLT & pLT)) & ~w.bnd;
  const uint32_t pm1 = (!(h.bnd >> 31) && pair_type(prev2, prev)) ? 1u : 0u;  // the byte before the window closes a pair
  const uint32_t overlap = pair2 & ((pair2 << 1) | pm1);                        // "<<=" ...: pairing is sequential there
  uint32_t nxpair = 0;
  if (nxb) {
    const uint32_t b31 = ((cLT >> 31) ? '<' : 0) | ((cGT >> 31) ? '>' : 0) | ((cBANG >> 31) ? '!' : 0);
    nxpair = pair_type((uint8_t)b31, (uint8_t)nxb) ? 1u : 0u;
  }
  const uint32_t first_of_pair = ((pair2 >> 1) | (nxpair << 31)) & cOP;
  const uint32_t has_op = cOP & ~first_of_pair;
  m.has = simple | has_word | has_dot | has_op | o.close;
  m.eofm = ((w.bnd >> 1) | ((uint32_t)(nx.bnd != 0) << 31)) & w.valid;
  m.TS = TSw | ((uint64_t)(simple | (cOP & ~pair2)) << 32) | (((pair2 & 1u) && !pm1) ? (1ull << 31) : 0ull);
  m.L64 = L64;
  m.DOT64 = DOT64;
  m.N64 = D64 | DOT64;
  // statements for the exact lexer.  End-of-token checks: an identifier must be followed by one of the characters of
  // tokenizer/mod.rs:486-503, a numeric literal (a digit-led run -- a plain hex literal has no such check, but one
  // followed by an odd character is rare enough for the exact lexer -- or "12.") by one of :506-543.
  const uint32_t okI = ((w.IE >> 1) | ((uint32_t)(nx.cls & 1u) << 31)) | nbnd;
  const uint32_t okN = ((w.NE >> 1) | ((uint32_t)((nx.cls >> 1) & 1u) << 31)) | nbnd;
  const uint32_t dledhi = (uint32_t)(dled >> 32);
  uint32_t bad = overlap | (cBANG & ~first_of_pair);
  bad |= E & ~dledhi & ~okI;
  bad |= ((has_word & dledhi) | (has_dot & (uint32_t)(joinL >> 32))) & ~okN;
  // a token that began before the look-back window
  if (((h.L | h.D | h.DOT) == 0xFFFFFFFFu) && !(w.bnd & 1u) && ((cL | cD | cDOT) & 1u)) bad |= 1u;
  m.bad = bad;
  m.bad_prev = 0;
}

// ---- token records -----------------------------------------------------------------------------------------
// word 0 of a record = first byte of the token (absolute); word 1 = end offset relative to the tile (1..8192) |
// kind << 14 | number of dots (saturated at 2) << 17 | a letter in the token << 19
enum : uint32_t {
  R3_GENERIC = 0,  // type follows from the first byte (words, one digit, punctuation, operators)
  R3_RAW = 1, R3_ESQ = 2, R3_EDQ = 3, R3_BT = 4, R3_EOF = 5, R3_FILL = 6,
  R3_NUM = 7       // digit- or dot-led token of more than one byte
};
#define NUTDB_R3_KIND_SHIFT 14
// flag bit of R3_ESQ / R3_EDQ records (shares the bit range of R3_NUM's dot count)
#define R3_STR_CLEAN 8u

// Rec: void operator()(uint32_t index, uint32_t start_abs, uint32_t end_abs, uint32_t flags /* kind | dots << 3 | letter << 5 */)
// index = index of the window's first token.
template <class Rec>
NUTDB_HD void win_records3(const WinCtx3& o, const TokMasks& m, uint32_t base, const StrCarry& sc_in, uint32_t index, Rec& rec) {
  const uint32_t TShi = (uint32_t)(m.TS >> 32), Nhi = (uint32_t)(m.N64 >> 32);
  uint32_t todo = m.has, n = 0;
  while (todo) {
    const int i = ctz32(todo);
    todo &= todo - 1;
    const uint32_t bit = 1u << i, end = base + (uint32_t)i + 1u;
    const uint32_t idx = index + n + (uint32_t)popc32(m.eofm & (bit - 1u));  // EOF tokens of statements that ended before
    n++;
    if (o.close & bit) {  // a literal / quoted identifier closes here
      const uint32_t opens = o.openm & (bit - 1u);
      uint32_t start;
      bool escd, chkd;
      if (opens) {
        start = base + (uint32_t)(31 - clz32(opens)) + 1u;
        escd = (o.escd & bit) != 0;
        chkd = (o.ustr & (bit - 1u) & ~((2u << (31 - clz32(opens))) - 1u)) != 0;  // between its two quotes
      } else {  // opened in an earlier window: offset and escaped flag come from the carry
        start = sc_in.open_pos + 1u;
        escd = (sc_in.esc | o.esc_first) != 0;
        chkd = sc_in.chk != 0 || (o.ustr & (bit - 1u)) != 0;
      }
      const uint32_t kind = (o.csq & bit) ? (escd ? (uint32_t)R3_ESQ : (uint32_t)R3_RAW)
                            : (o.cdq & bit) ? (escd ? (uint32_t)R3_EDQ : (uint32_t)R3_RAW) : (uint32_t)R3_BT;
      // (R3_STR_CLEAN: an escaped literal without a backslash-u escape unescapes without error -- literal.rs:45-102 can
      // only reject such an escape -- so the parser may skip its validation)
      rec(idx, start, end, kind | (chkd ? 0u : (uint32_t)R3_STR_CLEAN));
      continue;
    }
    const uint32_t incl = i >= 31 ? 0xFFFFFFFFu : ((2u << i) - 1u);
    const uint32_t ts = TShi & incl;
    if (ts) {  // the token starts in this window
      const int st = 31 - clz32(ts);
      uint32_t flags = R3_GENERIC;
      if (((Nhi >> st) & 1u) && st < i) {
        const uint32_t span = incl & ~((1u << st) - 1u);
        const int nd = popc32((uint32_t)(m.DOT64 >> 32) & span);
        flags = (uint32_t)R3_NUM | ((uint32_t)(nd > 2 ? 2 : nd) << 3) | (((uint32_t)(m.L64 >> 32) & span) ? 32u : 0u);
      }
      rec(idx, base + (uint32_t)st, end, flags);
    } else {  // ... in the previous one
      const uint32_t tl = (uint32_t)m.TS;
      const int st = tl ? 31 - clz32(tl) : 31;
      const uint64_t span = (((uint64_t)incl << 32) | 0xFFFFFFFFull) & ~((1ull << st) - 1ull);
      uint32_t flags = R3_GENERIC;
      if ((m.N64 >> st) & 1ull) {
        const int nd = popc64(m.DOT64 & span);
        flags = (uint32_t)R3_NUM | ((uint32_t)(nd > 2 ? 2 : nd) << 3) | ((m.L64 & span) ? 32u : 0u);
      }
      rec(idx, base - 32u + (uint32_t)st, end, flags);
    }
  }
  todo = m.eofm;  // the EOF token of a statement follows the token (if any) that ends at the statement's last byte
  while (todo) {
    const int i = ctz32(todo);
    todo &= todo - 1;
    const uint32_t bit = 1u << i, end = base + (uint32_t)i + 1u;
    rec(index + (uint32_t)popc32(m.has & (bit | (bit - 1u))) + (uint32_t)popc32(m.eofm & (bit - 1u)), end, end, (uint32_t)R3_EOF);
  }
}

// ---- one thread per token ----------------------------------------------------------------------------------
struct Tok3 {
  uint8_t type = NUTDB_TT_POISON, kw = 0, punt = 0;
  uint32_t start = 0, end = 0;  // statement relative
};
// Src: uint8_t at(abs) for bytes of the token itself (always staged), const uint8_t* span(abs, len).
// sst = first byte of the token's statement.
// Returns true for a word of 2..10 bytes: it may be a keyword, r.kw is still to be looked up (token_keyword3) -- the
// device does that in a second pass over the listed words so that all lanes of a warp run the hash together.
template <class Src>
NUTDB_HD bool token_finish3(const LexTables& T, Src& src, uint32_t start, uint32_t end, uint32_t flags, uint32_t sst, Tok3& r) {
  const uint32_t kind = flags & 7u, len = end - start;
  r.start = start - sst;
  r.end = end - sst;
  if (kind == R3_GENERIC) {
    const uint8_t b0 = src.at(start);
    uint8_t t = T.tt0[b0];
    if (t == NUTDB_TT_KeywordOrIdentifier) {  // tokenizer/mod.rs:262-282
      r.type = t;
      return len >= 2 && len <= 10;
    } else if (t == NUTDB_TT_IntegerLiteral) {
      r.kw = 1;  // one digit
    } else if (len == 2) {  // the only other generic tokens of two bytes: <= >= != <> << >>
      t = pair_type(b0, src.at(start + 1u));
      if (!t) t = NUTDB_TT_POISON;
    } else if (t == 0xFF) {
      t = NUTDB_TT_POISON;  // (a lone '!': its statement has been flagged by win_tokens3)
    }
    r.type = t;
  } else if (kind == R3_NUM) {
    const uint32_t ndots = (flags >> 3) & 3u;
    const uint8_t b0 = src.at(start);
    if (flags & 32u) {
      // "0" x|X hex-digits* (tokenizer/mod.rs:201-208): the span is the digits.  Only the plain form is lexed here;
      // "0x1G", "0x1.5", "1x", ".0x1" ... go to the exact lexer.
      bool ok = ndots == 0 && b0 == '0' && (src.at(start + 1u) | 0x20) == 'x';
      for (uint32_t q = start + 2u; ok && q < end; q++) ok = (T.prop[src.at(q)] & PR_HEX) != 0;
      if (!ok) { r.punt = 1; return false; }
      r.type = NUTDB_TT_HexLiteral;
      r.start += 2u;
      r.kw = (uint8_t)(len - 2u > 255u ? 255u : len - 2u);
    } else if (ndots == 0) {  // tokenizer/mod.rs:196-238
      r.type = NUTDB_TT_IntegerLiteral;
      r.kw = (uint8_t)(len > 255u ? 255u : len);
    } else if (ndots == 1) {  // digits '.' digits* | '.' digits (tokenizer/mod.rs:246-258)
      r.type = NUTDB_TT_FloatLiteral;
    } else {
      r.punt = 1;
    }
  } else if (kind == R3_EOF) {
    r.type = NUTDB_TT_EOF;
    r.start = r.end = end - sst;
  } else if (kind == R3_FILL) {
    r.type = NUTDB_TT_POISON;
    r.start = r.end = 0;
  } else {  // payload between the quotes
    r.type = kind == R3_RAW ? (uint8_t)NUTDB_TT_RawStringLiteral
             : kind == R3_ESQ ? (uint8_t)NUTDB_TT_EscapedSQStringLiteral
             : kind == R3_EDQ ? (uint8_t)NUTDB_TT_EscapedDQStringLiteral : (uint8_t)NUTDB_TT_DelimitedIdentifier;
    r.end = end - 1u - sst;
    if ((kind == R3_ESQ || kind == R3_EDQ) && (flags & R3_STR_CLEAN)) r.kw = 1;  // tok_kw of an escaped literal: 1 = no \u escape
  }
  return false;
}
// keyword id of a word of 2..10 bytes (the shared-memory perfect hash)
template <class Src>
NUTDB_HD uint8_t token_keyword3(const LexTables& T, Src& src, uint32_t start, uint32_t len) {
  const uint8_t* wp = src.span(start, len);
  if (wp) {
    uint32_t w0, w1, w2;
    load_word12(wp, len, w0, w1, w2);
    return keyword_lookup_words(T, len, w0, w1, w2);
  }
  Src& sr = src;
  return keyword_lookup(T, len, [&sr, start](uint32_t q) { return sr.at(start + q); });
}

}  // namespace nlex3
