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
  w.bt = r6 & h00 & l00;
  w.D = r3 & n7 & (~b3 | (~b2 & ~b1));
  const uint32_t lo_nz = b3 | b2 | b1 | b0, lo_leA = ~b3 | (~b2 & ~l11);
  w.L = (n7 & (((r4 | r6) & lo_nz) | ((r5 | r7) & lo_leA))) | (r5 & h11 & l11);
  // single-character tokens: % & ( ) * + ,   : ;   [ ] ^   { | } ~
  w.P = (r2 & ((h01 & (b1 ^ b0)) | h10 | (h11 & l00))) | (r3 & h10 & b1) | (r5 & ((h10 & l11) | (h11 & (b1 ^ b0)))) |
        (r7 & ((h10 & l11) | (h11 & ~l11)));
  w.OP = op.LT | op.GT | op.EQ | op.BANG;
  w.IE = 0;
  w.NE = 0;
  w.valid = valid;
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
  uint64_t L64 = 0, DOT64 = 0;
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
NUTDB_HD void win_tokens3(const Win& w, const Ops& op, const WinCtx& o, const Hist3& h, const Next& nx, uint8_t prev,
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
  // statements for the exact lexer, as far as masks can tell (token_finish3 adds the end-of-token checks)
  uint32_t bad = overlap | (cBANG & ~first_of_pair);
  // a token that began before the look-back window
  if (((h.L | h.D | h.DOT) == 0xFFFFFFFFu) && !(w.bnd & 1u) && ((cL | cD | cDOT) & 1u)) bad |= 1u;
  m.bad = bad;
  m.bad_prev = 0;
}

// ---- token records -----------------------------------------------------------------------------------------
// flags of a record: [0:2) number of dots in the token (saturated), [2] a letter in the token, [3:6) explicit type
enum : uint32_t { R3_CODE = 0, R3_RAW = 1, R3_ESQ = 2, R3_EDQ = 3, R3_BT = 4, R3_EOF = 5, R3_FILL = 6 };
#define NUTDB_R3_FLAG_SHIFT 14   // word 1 of a record = end offset relative to the tile (1..8192) | flags << 14

// Rec: void operator()(uint32_t index, uint32_t start_abs, uint32_t end_abs, uint32_t flags)
// index = index of the window's first token; returns the index after its last one.
template <class Rec>
NUTDB_HD uint32_t win_records3(const WinCtx& o, const TokMasks& m, uint32_t base, const StrCarry& sc_in, uint32_t index, Rec& rec) {
  uint32_t todo = m.has | m.eofm, c = 0;
  while (todo) {
    const int i = ctz32(todo);
    todo &= todo - 1;
    const uint32_t bit = 1u << i, end = base + (uint32_t)i + 1u;
    if (m.has & bit) {
      if (o.close & bit) {
        if (c < o.ncap) {  // a literal / quoted identifier closes here (ctx_window recorded it)
          const uint32_t cw = (o.capw[c >> 1] >> (16u * (c & 1u))) & 0xFFFFu;
          const uint32_t code = (cw >> 11) & 3u;
          uint32_t start = base + ((cw >> 5) & 31u) + 1u, kind;
          if ((cw >> 10) & 1u) {  // opened in an earlier window: offset and escaped flag come from the carry
            start = sc_in.open_pos + 1u;
            const bool escd = (sc_in.esc | o.esc_first) != 0;
            kind = code == 3 ? (uint32_t)R3_BT : (escd ? (code == 1 ? (uint32_t)R3_ESQ : (uint32_t)R3_EDQ) : (uint32_t)R3_RAW);
          } else {
            kind = code == 0 ? (uint32_t)R3_RAW : code == 1 ? (uint32_t)R3_ESQ : code == 2 ? (uint32_t)R3_EDQ : (uint32_t)R3_BT;
          }
          rec(index, start, end, kind << 3);
          c++;
        } else {  // beyond the capture array: the statement is flagged; a fixed filler keeps the arrays reproducible
          rec(index, 0u, end, (uint32_t)R3_FILL << 3);
        }
      } else {
        const int p = 32 + i;
        const uint64_t below = p >= 63 ? ~0ull : ((2ull << p) - 1ull);
        const uint64_t ts = m.TS & below;
        const int st = ts ? 63 - clz64(ts) : p;
        const uint64_t span = below & ~((1ull << st) - 1ull);
        const int nd = popc64(m.DOT64 & span);
        rec(index, base + (uint32_t)st - 32u, end, (uint32_t)(nd > 2 ? 2 : nd) | ((m.L64 & span) ? 4u : 0u));
      }
      index++;
    }
    if (m.eofm & bit) {
      rec(index, end, end, (uint32_t)R3_EOF << 3);
      index++;
    }
  }
  return index;
}

// ---- one thread per token ----------------------------------------------------------------------------------
struct Tok3 {
  uint8_t type = NUTDB_TT_POISON, kw = 0, punt = 0;
  uint32_t start = 0, end = 0;  // statement relative
};
// Src: uint8_t byte(abs), const uint8_t* span(abs, len).  sst = first byte of the token's statement; next_bnd = the
// byte after the token starts another statement (or is the end of the batch).
template <class Src>
NUTDB_HD void token_finish3(const LexTables& T, Src& src, uint32_t start, uint32_t end, uint32_t flags, uint32_t sst,
                            bool next_bnd, Tok3& r) {
  const uint32_t kind = (flags >> 3) & 7u;
  if (kind != R3_CODE) {
    if (kind == R3_EOF) {
      r.type = NUTDB_TT_EOF;
      r.start = r.end = end - sst;
    } else if (kind == R3_FILL) {
      r.type = NUTDB_TT_POISON;
    } else {  // payload between the quotes
      r.type = kind == R3_RAW ? (uint8_t)NUTDB_TT_RawStringLiteral
               : kind == R3_ESQ ? (uint8_t)NUTDB_TT_EscapedSQStringLiteral
               : kind == R3_EDQ ? (uint8_t)NUTDB_TT_EscapedDQStringLiteral : (uint8_t)NUTDB_TT_DelimitedIdentifier;
      r.start = start - sst;
      r.end = end - 1u - sst;
    }
    return;
  }
  const uint32_t len = end - start, ndots = flags & 3u;
  const bool letter = (flags & 4u) != 0;
  const uint8_t b0 = src.byte(start);
  const uint8_t nb = next_bnd ? (uint8_t)0 : src.byte(end);
  const bool end_ident = next_bnd || (T.prop[nb] & PR_IDENT_END);  // tokenizer/mod.rs:486-503
  const bool end_num = next_bnd || (T.prop[nb] & PR_NUM_END);      // tokenizer/mod.rs:506-543
  r.start = start - sst;
  r.end = end - sst;
  const uint8_t pr = T.prop[b0];
  if (pr & PR_DIGIT) {
    if (letter) {
      // "0" x|X hex-digits* (tokenizer/mod.rs:201-208): the span is the digits, there is no end-of-token check.
      // Only the plain form is lexed here; "0x1G", "0x1.5", "1x" ... go to the exact lexer.
      bool ok = ndots == 0 && len >= 2 && b0 == '0' && (src.byte(start + 1u) | 0x20) == 'x';
      for (uint32_t q = start + 2u; ok && q < end; q++) ok = (T.prop[src.byte(q)] & PR_HEX) != 0;
      if (!ok) { r.punt = 1; return; }
      r.type = NUTDB_TT_HexLiteral;
      r.start += 2u;
      r.kw = (uint8_t)(len - 2u > 255u ? 255u : len - 2u);
    } else if (ndots == 0) {  // tokenizer/mod.rs:196-238
      if (!end_num) { r.punt = 1; return; }
      r.type = NUTDB_TT_IntegerLiteral;
      r.kw = (uint8_t)(len > 255u ? 255u : len);
    } else {  // digits '.' digits* (tokenizer/mod.rs:246-258)
      if (ndots != 1 || !end_num) { r.punt = 1; return; }
      r.type = NUTDB_TT_FloatLiteral;
    }
  } else if (pr & PR_WORD) {  // identifier / keyword (tokenizer/mod.rs:262-282)
    if (!end_ident) { r.punt = 1; return; }
    r.type = NUTDB_TT_KeywordOrIdentifier;
    if (len >= 2 && len <= 10) {
      const uint8_t* wp = src.span(start, len);
      if (wp) {
        uint32_t w0, w1, w2;
        load_word12(wp, len, w0, w1, w2);
        r.kw = keyword_lookup_words(T, len, w0, w1, w2);
      } else {
        Src& sr = src;
        r.kw = keyword_lookup(T, len, [&sr, start](uint32_t q) { return sr.byte(start + q); });
      }
    }
  } else if (b0 == '.') {
    if (len == 1) {
      r.type = NUTDB_TT_Dot;  // no end check (tokenizer/mod.rs:248-250)
    } else {
      if (letter || ndots != 1 || !end_num) { r.punt = 1; return; }
      r.type = NUTDB_TT_FloatLiteral;
    }
  } else if (len == 1) {
    const uint8_t t = b0 == '-' ? (uint8_t)NUTDB_TT_Minus
                      : b0 == '/' ? (uint8_t)NUTDB_TT_Div
                      : b0 == '<' ? (uint8_t)NUTDB_TT_Lt
                      : b0 == '>' ? (uint8_t)NUTDB_TT_Gt
                      : b0 == '=' ? (uint8_t)NUTDB_TT_Eq : T.single_tt[b0];
    if (t == 0xFF) { r.punt = 1; return; }
    r.type = t;
  } else {
    const uint8_t t = len == 2 ? pair_type(b0, src.byte(start + 1u)) : (uint8_t)0;
    if (!t) { r.punt = 1; return; }
    r.type = t;
  }
}

}  // namespace nlex3
