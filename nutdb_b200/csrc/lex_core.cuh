// lex_core.cuh -- the lexer as a pair of small finite automata plus a resumable token walker.
//
// Replaces the reference's sequential tokenizer (src/parser/tokenizer/mod.rs:66-543 on top of
// Utf8Iter, tokenizer/utf8_iter.rs:126-237) with a formulation that can start at ANY byte once
// two 3-bit states are known there, so the batch can be cut into fixed 32-byte chunks:
//
//   automaton A ("context", 8 states): code / '..' / ".." / `..` / --comment / block comment.
//       Its input per byte is an EVENT computed from (byte, previous byte, backslash parity).
//       Quotes toggle code<->string (a doubled quote '' is two toggles, which leaves the context
//       exactly where the reference's "peek next quote" logic leaves it, mod.rs:131-145).
//   automaton B ("code token", 8 states): which multi-byte code token we are inside
//       (word / 0 / int / . / frac / hex / first char of a 2-char operator).  It only runs on
//       bytes that A says are code, and resets otherwise.
//
// Both have 8 states, so a transition FUNCTION is 8 nibbles = one 32-bit word and composing two
// functions is two PRMT byte-permutes on sm_100a (vec8_* below).  The kernel computes each
// chunk's function, prefix-scans them (warp shuffles + decoupled look-back across tiles) and
// then re-walks every chunk from its now-known entry state with lex_walk(), which emits tokens.
//
// Everything here is NUTDB_HD so tests can drive the identical code on the host.
#pragma once
#include <stdint.h>

#include "../../include/nutdb_gpu.h"

#if defined(__CUDACC__)
#define NUTDB_HD __host__ __device__ __forceinline__
#define NUTDB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define NUTDB_HD inline
#define NUTDB_HD_NOINLINE inline
#endif

namespace nlex {

// ------------------------------------------------------------------------------------------
// automaton A: context
// ------------------------------------------------------------------------------------------
enum : uint8_t { A_C = 0, A_CX = 1, A_SQ = 2, A_DQ = 3, A_BT = 4, A_LC = 5, A_BC0 = 6, A_BC = 7 };
// A_CX  = code, but the previous byte is the '/' that closed a block comment (so "/*" cannot start here)
// A_BC0 = block comment, previous byte is the '*' of the opener (so "*/" cannot close here: "/*/" is open)
enum : uint8_t { EV_OTHER = 0, EV_SQ, EV_DQ, EV_BT, EV_DD, EV_NL, EV_SLST, EV_STSL, EV_COUNT };

// esc = the byte is preceded by an odd run of backslashes (only meaningful inside '..' / "..",
// where the reference consumes "\\" plus the next char, mod.rs:147-158)
NUTDB_HD uint8_t a_event(uint8_t b, uint8_t prev, bool esc) {
  switch (b) {
    case '\'': return esc ? EV_OTHER : EV_SQ;
    case '"': return esc ? EV_OTHER : EV_DQ;
    case '`': return EV_BT;
    case '-': return prev == '-' ? EV_DD : EV_OTHER;
    case '\n':
    case '\r': return EV_NL;
    case '*': return prev == '/' ? EV_SLST : EV_OTHER;
    case '/': return prev == '*' ? EV_STSL : EV_OTHER;
    default: return EV_OTHER;
  }
}

NUTDB_HD uint8_t a_next(uint8_t s, uint8_t ev) {
  switch (s) {
    case A_C:
      switch (ev) {
        case EV_SQ: return A_SQ;
        case EV_DQ: return A_DQ;
        case EV_BT: return A_BT;
        case EV_DD: return A_LC;     // "--"  mod.rs:371-375
        case EV_SLST: return A_BC0;  // "/*"  mod.rs:384-388
        default: return A_C;
      }
    case A_CX:
      switch (ev) {
        case EV_SQ: return A_SQ;
        case EV_DQ: return A_DQ;
        case EV_BT: return A_BT;
        default: return A_C;  // incl. EV_SLST: the '/' belonged to the closed comment, '*' is Mul
      }
    case A_SQ: return ev == EV_SQ ? A_C : A_SQ;
    case A_DQ: return ev == EV_DQ ? A_C : A_DQ;
    case A_BT: return ev == EV_BT ? A_C : A_BT;
    case A_LC: return ev == EV_NL ? A_C : A_LC;  // mod.rs:434
    case A_BC0: return A_BC;
    default: return ev == EV_STSL ? A_CX : A_BC;  // A_BC, mod.rs:446-466
  }
}

// ------------------------------------------------------------------------------------------
// automaton B: code tokens
// ------------------------------------------------------------------------------------------
enum : uint8_t { B_N = 0, B_W = 1, B_Z = 2, B_I = 3, B_D0 = 4, B_F = 5, B_H = 6, B_OP = 7 };
enum : uint8_t {
  CL_RESET = 0,  // byte is not code (inside string / comment / closing delimiter)
  CL_WS, CL_PUNCT, CL_LT, CL_GT, CL_BANG, CL_EQ, CL_LT_P, CL_GT_P, CL_EQ_P,
  CL_ZERO, CL_DIG, CL_HEXL, CL_XL, CL_LET, CL_DOT, CL_AT, CL_DOL, CL_BAD, CL_COUNT
};

// byte property bits
enum : uint8_t {
  PR_IDENT_END = 1,  // valid char after identifier / config identifier  (tokenizer/mod.rs:486-503)
  PR_NUM_END = 2,    // valid char after numeric / query parameter       (tokenizer/mod.rs:506-543)
  PR_WORD = 4,       // [A-Za-z0-9_]
  PR_DIGIT = 8,
  PR_HEX = 16,
  PR_WS = 32
};

NUTDB_HD uint8_t b_next(uint8_t s, uint8_t cl) {
  if (cl == CL_RESET) return B_N;
  const bool digit = cl == CL_ZERO || cl == CL_DIG;
  const bool letter = cl == CL_HEXL || cl == CL_XL || cl == CL_LET;
  switch (s) {
    case B_W:
      if (digit || letter) return B_W;
      break;
    case B_Z:
      if (digit) return B_I;
      if (cl == CL_XL) return B_H;
      if (cl == CL_DOT) return B_F;
      break;
    case B_I:
      if (digit) return B_I;
      if (cl == CL_DOT) return B_F;
      break;
    case B_D0:
      if (digit) return B_F;
      break;
    case B_F:
      if (digit) return B_F;
      break;
    case B_H:
      if (digit || cl == CL_HEXL) return B_H;
      break;
    case B_OP:
      if (cl == CL_LT_P || cl == CL_GT_P || cl == CL_EQ_P) return B_N;
      break;
    default: break;
  }
  // the byte starts fresh
  switch (cl) {
    case CL_LT: case CL_GT: case CL_BANG: case CL_LT_P: case CL_GT_P: return B_OP;
    case CL_ZERO: return B_Z;
    case CL_DIG: case CL_DOL: return B_I;
    case CL_HEXL: case CL_XL: case CL_LET: case CL_AT: return B_W;
    case CL_DOT: return B_D0;
    default: return B_N;
  }
}

// ------------------------------------------------------------------------------------------
// tables (built once on the host, copied to the device, staged in shared memory per CTA)
// ------------------------------------------------------------------------------------------
#define NUTDB_KW_SLOTS 512
struct LexTables {
  uint8_t base_cls[256];   // class ignoring the pair rule
  uint8_t prop[256];
  uint8_t single_tt[256];  // token type of single-char tokens, 0xFF otherwise
  uint8_t tt0[256];        // token type by FIRST byte: words 0, digits Integer, '.', '-', '/', '<', '>', '=', single_tt; else 0xFF
  uint32_t a_row[EV_COUNT][2];   // transition function of each event, one state per BYTE (lo: states 0-3, hi: 4-7)
  uint32_t b_row[CL_COUNT][2];
  uint32_t kw_mul[4];            // perfect hash multipliers
  uint8_t kw_slot[NUTDB_KW_SLOTS];  // keyword id or 0
  uint8_t kw_len[NUTDB_KW_COUNT + 1];
  alignas(4) uint8_t kw_text[NUTDB_KW_COUNT + 1][12];  // lower case ([a-z0-9] only), zero padded; read as 3 words
};

NUTDB_HD uint8_t b_class(const LexTables& T, uint8_t b, uint8_t prev) {
  uint8_t cl = T.base_cls[b];
  if (cl == CL_LT && prev == '<') return CL_LT_P;                                    // "<<"
  if (cl == CL_GT && (prev == '<' || prev == '>')) return CL_GT_P;                   // "<>" ">>"
  if (cl == CL_EQ && (prev == '<' || prev == '>' || prev == '!')) return CL_EQ_P;    // "<=" ">=" "!="
  return cl;
}

// ------------------------------------------------------------------------------------------
// 8-state transition functions packed as 8 nibbles (nibble s = image of state s)
// ------------------------------------------------------------------------------------------
#define NUTDB_VEC8_ID 0x76543210u

NUTDB_HD uint32_t vec8_apply(uint32_t f, uint32_t s) { return (f >> (4 * s)) & 7u; }

NUTDB_HD uint32_t vec8_pack_bytes(uint32_t lo, uint32_t hi) {  // 8 bytes (values 0..7) -> 8 nibbles
#if defined(__CUDA_ARCH__)
  uint32_t l = __byte_perm(lo | (lo >> 4), 0, 0x4420);
  uint32_t h = __byte_perm(hi | (hi >> 4), 0, 0x4420);
  return l | (h << 16);
#else
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) {
    r |= ((lo >> (8 * i)) & 7u) << (4 * i);
    r |= ((hi >> (8 * i)) & 7u) << (4 * (i + 4));
  }
  return r;
#endif
}

// run' = row o run   (first run, then the byte whose function is `row`, given one state per byte)
NUTDB_HD uint32_t vec8_then_row(uint32_t run, uint32_t row_lo, uint32_t row_hi) {
#if defined(__CUDA_ARCH__)
  uint32_t lo = __byte_perm(row_lo, row_hi, run & 0xFFFFu);
  uint32_t hi = __byte_perm(row_lo, row_hi, run >> 16);
  return vec8_pack_bytes(lo, hi);
#else
  uint32_t r = 0;
  for (int s = 0; s < 8; s++) {
    uint32_t m = (run >> (4 * s)) & 7u;
    uint32_t img = m < 4 ? (row_lo >> (8 * m)) & 7u : (row_hi >> (8 * (m - 4))) & 7u;
    r |= img << (4 * s);
  }
  return r;
#endif
}

NUTDB_HD void vec8_unpack(uint32_t g, uint32_t& lo, uint32_t& hi) {  // nibbles -> bytes
  uint32_t x = g & 0xFFFFu, y = g >> 16;
  lo = (x & 0xFu) | ((x & 0xF0u) << 4) | ((x & 0xF00u) << 8) | ((x & 0xF000u) << 12);
  hi = (y & 0xFu) | ((y & 0xF0u) << 4) | ((y & 0xF00u) << 8) | ((y & 0xF000u) << 12);
}

// (g o f): first f, then g
NUTDB_HD uint32_t vec8_then(uint32_t f, uint32_t g) {
  uint32_t lo, hi;
  vec8_unpack(g, lo, hi);
  return vec8_then_row(f, lo, hi);
}

// ------------------------------------------------------------------------------------------
// keyword perfect hash (the "shared-memory perfect hash" of stage 2).  `get(i)` returns byte i
// of the word.  Words are case-folded exactly like eq_ignore_ascii_case (mod.rs:53-57).
// ------------------------------------------------------------------------------------------
NUTDB_HD uint8_t ascii_lower(uint8_t c) { return (c >= 'A' && c <= 'Z') ? (uint8_t)(c | 0x20) : c; }

template <class Get>
NUTDB_HD uint8_t keyword_lookup(const LexTables& T, uint32_t len, Get get) {
  if (len < 2 || len > 10) return 0;
  uint32_t c0 = ascii_lower(get(0)), c1 = ascii_lower(get(1)), cl = ascii_lower(get(len - 1)),
           cp = ascii_lower(get(len - 2));
  uint32_t h = (c0 * T.kw_mul[0] + c1 * T.kw_mul[1] + cl * T.kw_mul[2] + cp * T.kw_mul[3] + len) & (NUTDB_KW_SLOTS - 1);
  uint8_t id = T.kw_slot[h];
  if (id == 0 || T.kw_len[id] != len) return 0;
  for (uint32_t i = 0; i < len; i++)
    if (ascii_lower(get(i)) != T.kw_text[id][i]) return 0;
  return id;
}

// The same look-up on the word held in registers: w0..w2 are its first 12 bytes (little endian; bytes at and
// beyond `len` are arbitrary).  Folding with `| 0x20` instead of ascii_lower is exact here: the word consists
// of [A-Za-z0-9_] and a keyword of [a-z0-9], so c | 0x20 == k exactly when c equals k ignoring ASCII case
// ('_' becomes 0x7F, which no keyword contains).
NUTDB_HD uint8_t keyword_lookup_words(const LexTables& T, uint32_t len, uint32_t w0, uint32_t w1, uint32_t w2) {
  if (len < 2 || len > 10) return 0;
  const uint32_t L0 = w0 | 0x20202020u, L1 = w1 | 0x20202020u, L2 = w2 | 0x20202020u;
  const uint64_t lo = (uint64_t)L0 | ((uint64_t)L1 << 32);
  const uint32_t il = len - 1, ip = len - 2;
  const uint32_t c0 = L0 & 255u, c1 = (L0 >> 8) & 255u;
  const uint32_t cl = il < 8 ? (uint32_t)(lo >> (8 * il)) & 255u : (L2 >> (8 * (il - 8))) & 255u;
  const uint32_t cp = ip < 8 ? (uint32_t)(lo >> (8 * ip)) & 255u : (L2 >> (8 * (ip - 8))) & 255u;
  const uint32_t h = (c0 * T.kw_mul[0] + c1 * T.kw_mul[1] + cl * T.kw_mul[2] + cp * T.kw_mul[3] + len) & (NUTDB_KW_SLOTS - 1);
  const uint8_t id = T.kw_slot[h];
  if (id == 0 || T.kw_len[id] != len) return 0;
  const uint32_t m0 = len >= 4 ? 0xFFFFFFFFu : ((1u << (8 * len)) - 1u);
  const uint32_t m1 = len >= 8 ? 0xFFFFFFFFu : (len <= 4 ? 0u : ((1u << (8 * (len - 4))) - 1u));
  const uint32_t m2 = len <= 8 ? 0u : ((1u << (8 * (len - 8))) - 1u);
  const uint32_t* k = reinterpret_cast<const uint32_t*>(T.kw_text[id]);
  return ((L0 & m0) == k[0] && (L1 & m1) == k[1] && (L2 & m2) == k[2]) ? id : (uint8_t)0;
}
// the first 12 bytes of a word that lies contiguously in memory at p (device: shared memory, read as aligned
// words -- up to 15 bytes around the word are touched, always inside the staged tile and its neighbours)
NUTDB_HD void load_word12(const uint8_t* p, uint32_t len, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
#ifdef __CUDA_ARCH__
  const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
  const uint32_t* q = reinterpret_cast<const uint32_t*>(p - a);
  const uint32_t x0 = q[0], x1 = q[1], x2 = q[2], x3 = q[3];
  const uint32_t sh = 8u * a;
  w0 = __funnelshift_r(x0, x1, sh);
  w1 = __funnelshift_r(x1, x2, sh);
  w2 = __funnelshift_r(x2, x3, sh);
#else
  uint32_t w[3] = {0, 0, 0};
  for (uint32_t i = 0; i < len && i < 12; i++) w[i >> 2] |= (uint32_t)p[i] << (8 * (i & 3));
  w0 = w[0];
  w1 = w[1];
  w2 = w[2];
#endif
}

// ------------------------------------------------------------------------------------------
// resumable walker
// ------------------------------------------------------------------------------------------
enum : uint8_t {
  PK_NONE = 0, PK_WS, PK_IDENT, PK_CFG, PK_QP, PK_NUM, PK_OP, PK_DASH, PK_SLASH, PK_SQ, PK_DQ, PK_BT, PK_LC, PK_BC
};

// Everything a chunk needs to know about the bytes before it.
struct LexCarry {
  uint8_t A = A_C, B = B_N;
  uint8_t pk = PK_NONE;       // kind of the token in progress
  uint8_t escaped = 0;        // string in progress has seen '' / "" / backslash (mod.rs:122,137,157)
  uint8_t esc = 0;            // next byte is preceded by an odd backslash run
  uint8_t prev_esc = 0;       // esc of the previous byte (for "\\\r\n", mod.rs:152-156)
  uint8_t prev = 0;           // previous byte of the same statement, 0 at statement start
  uint32_t tok_start = 0;     // absolute offset of the first byte of the token in progress
  uint32_t stmt_start = 0;    // absolute offset of the current statement
  uint32_t count = 0;         // tokens emitted so far (index of the next token)
  uint32_t seg = 0;           // index of the current (non-empty) statement
  uint32_t nseg_seen = 0;     // statement starts processed by this walker
};

// Sink concept:
//   void token(uint32_t index, uint8_t type, uint32_t start_rel, uint32_t end_rel, uint8_t kw);
//   void seg_begin(uint32_t seg, uint32_t first_token_index, uint32_t stmt_start_abs);
//   void seg_end(uint32_t seg, uint32_t end_token_index, uint32_t stmt_start_abs);
// Src concept:
//   uint8_t byte(uint32_t abs_pos);       // any position of the batch (used for short look-backs)
//   bool boundary(uint32_t abs_pos);      // a statement starts at abs_pos

template <bool EmitAll, class Src, class Sink>
struct Walker {
  const LexTables& T;
  Src& src;
  Sink& sink;
  LexCarry c;
  bool counting = false;  // phase C: only the NUMBER of tokens matters; tok_start may be unknown

  NUTDB_HD Walker(const LexTables& t, Src& s, Sink& k, const LexCarry& carry) : T(t), src(s), sink(k), c(carry) {}

  NUTDB_HD void emit(uint8_t type, uint32_t start_abs, uint32_t end_abs, uint8_t kw = 0) {
    if (!EmitAll && (type == NUTDB_TT_Whitespace || type == NUTDB_TT_Comment)) return;
    // side byte of integer / hex literals: number of digits (saturated), so the parser can accept the
    // common "too short to overflow" case (literal.rs:18-31) without touching the text
    if (type == NUTDB_TT_IntegerLiteral || type == NUTDB_TT_HexLiteral)
      kw = (uint8_t)((end_abs - start_abs) > 255u ? 255u : (end_abs - start_abs));
    if (!counting) sink.token(c.count, type, start_abs - c.stmt_start, end_abs - c.stmt_start, kw);
    c.count++;
  }
  NUTDB_HD void poison(uint32_t pos_abs, uint32_t site) {
    if (!counting) sink.token(c.count, NUTDB_TT_POISON, pos_abs - c.stmt_start, site, 0);
    c.count++;
  }

  NUTDB_HD void emit_ident(uint32_t s, uint32_t e) {
    uint32_t len = e - s;
    uint8_t kw = 0;
    if (!counting && len >= 2 && len <= 10) {
      Src& sr = src;
      kw = keyword_lookup(T, len, [&sr, s](uint32_t i) { return sr.byte(s + i); });
    }
    emit(NUTDB_TT_KeywordOrIdentifier, s, e, kw);
  }

  // side byte of an escaped string literal: 1 = it holds no backslash-u escape, so unescaping it (literal.rs:45-102)
  // cannot fail and the parser need not look at its bytes; the mask lexer computes the same bit (R3_STR_CLEAN)
  NUTDB_HD bool no_escaped_u(uint32_t s, uint32_t e) {
    for (uint32_t p = s; p < e; p++)
      if (src.byte(p) == '\\') {
        if (++p < e && src.byte(p) == 'u') return false;
      }
    return true;
  }

  // block comment payload end: start of the star run before the closing '/', not before the body
  // (the reference keeps `end` at the last position where comment_end == 0, mod.rs:441-460)
  NUTDB_HD uint32_t bc_payload_end(uint32_t slash_pos) {
    uint32_t e = slash_pos;
    uint32_t body = c.tok_start + 2;
    while (e > body && src.byte(e - 1) == '*') e--;
    return e;
  }
  NUTDB_HD uint32_t lc_payload_start(uint32_t end_pos) {
    uint32_t s = c.tok_start + 2;
    while (s < end_pos && src.byte(s) == ' ') s++;  // skip(' ')  mod.rs:432
    return s;
  }
  NUTDB_HD void emit_comment_lc(uint32_t end_pos) {
    if (!EmitAll) return;
    if (counting) { c.count++; return; }
    emit(NUTDB_TT_Comment, lc_payload_start(end_pos), end_pos);
  }
  NUTDB_HD void emit_comment_bc(uint32_t slash_pos) {
    if (!EmitAll) return;
    if (counting) { c.count++; return; }
    emit(NUTDB_TT_Comment, c.tok_start + 2, bc_payload_end(slash_pos));
  }

  // The token in progress ends before `pos`: the byte at pos (or EOF) does not continue it.
  // Emits EXACTLY ONE token (or poison) for every pk except PK_NONE, whatever the first byte of
  // the token turns out to be -- phase C relies on that to count without knowing tok_start.
  NUTDB_HD void finish_code_token(uint32_t pos, bool ident_end_ok, bool num_end_ok) {
    uint8_t pk = c.pk;
    c.pk = PK_NONE;
    if (pk == PK_NONE) return;
    if (counting) {
      if (EmitAll || pk != PK_WS) c.count++;
      return;
    }
    const uint32_t ts = c.tok_start;
    if (pk == PK_IDENT && src.byte(ts) == '@') pk = PK_CFG;
    if (pk == PK_NUM && src.byte(ts) == '$') pk = PK_QP;
    switch (pk) {
      case PK_WS: emit(NUTDB_TT_Whitespace, ts, pos); break;
      case PK_IDENT:
        if (!ident_end_ok) poison(pos, NUTDB_LE_IDENT_END);
        else emit_ident(ts, pos);
        break;
      case PK_CFG:
        if (pos > ts + 1 && (T.prop[src.byte(ts + 1)] & PR_DIGIT)) poison(ts + 1, NUTDB_LE_CFG_DIGIT);  // mod.rs:290
        else if (!ident_end_ok) poison(pos, NUTDB_LE_CFG_END);
        else if (pos == ts + 1) poison(pos, NUTDB_LE_CFG_EMPTY);
        else emit(NUTDB_TT_ConfigIdentifier, ts + 1, pos);
        break;
      case PK_QP: {
        // digits only; automaton B lets a '.' through (B_I -> B_F), which the reference rejects
        uint32_t q = ts + 1;
        while (q < pos && (T.prop[src.byte(q)] & PR_DIGIT)) q++;
        if (q < pos) poison(q, NUTDB_LE_QP_END);
        else if (!num_end_ok) poison(pos, NUTDB_LE_QP_END);
        else if (pos == ts + 1) poison(pos, NUTDB_LE_QP_EMPTY);
        else emit(NUTDB_TT_QueryParameter, ts + 1, pos);
        break;
      }
      case PK_NUM:
        switch (c.B) {
          case B_Z:
            if (!num_end_ok) poison(pos, NUTDB_LE_NUM_ZERO);
            else emit(NUTDB_TT_IntegerLiteral, ts, pos);
            break;
          case B_I:
            if (!num_end_ok) poison(pos, NUTDB_LE_NUM_INT);
            else emit(NUTDB_TT_IntegerLiteral, ts, pos);
            break;
          case B_D0: emit(NUTDB_TT_Dot, ts, pos); break;  // no end check, mod.rs:248-250
          case B_F:
            if (!num_end_ok) poison(pos, NUTDB_LE_NUM_FLOAT);
            else emit(NUTDB_TT_FloatLiteral, ts, pos);
            break;
          case B_H: emit(NUTDB_TT_HexLiteral, ts + 2, pos); break;  // no end check, mod.rs:201-208
          default: poison(pos, NUTDB_LE_INVALID_CHAR); break;     // only reachable after an earlier poison
        }
        break;
      case PK_OP:
        if (c.prev == '!') poison(pos, NUTDB_LE_BANG);
        else emit(c.prev == '<' ? NUTDB_TT_Lt : NUTDB_TT_Gt, ts, pos);
        break;
      case PK_DASH: emit(NUTDB_TT_Minus, ts, pos); break;
      case PK_SLASH: emit(NUTDB_TT_Div, ts, pos); break;
      case PK_SQ:
      case PK_DQ:  // closed by the quote at pos-1 (we are in code context again)
        if (c.escaped)
          emit(pk == PK_SQ ? NUTDB_TT_EscapedSQStringLiteral : NUTDB_TT_EscapedDQStringLiteral, ts + 1, pos - 1,
               counting ? (uint8_t)0 : (uint8_t)no_escaped_u(ts + 1, pos - 1));
        else
          emit(NUTDB_TT_RawStringLiteral, ts + 1, pos - 1);
        break;
      default: poison(pos, NUTDB_LE_INVALID_CHAR); break;  // PK_BT/LC/BC in code context: only after a poison
    }
  }

  // End of the statement at absolute offset e.
  NUTDB_HD void flush_eof(uint32_t e) {
    switch (c.A) {
      case A_SQ:
      case A_DQ: poison(e, NUTDB_LE_STR_EOF); break;
      case A_BT: poison(e, e == c.tok_start + 1 ? NUTDB_LE_BT_EMPTY : NUTDB_LE_BT_EOF); break;
      case A_LC: emit_comment_lc(e); break;
      case A_BC0:
      case A_BC: poison(e, NUTDB_LE_BC_EOF); break;
      default: finish_code_token(e, true, true); break;
    }
    emit(NUTDB_TT_EOF, e, e);
    if (!counting) sink.seg_end(c.seg, c.count, c.stmt_start);
  }

  NUTDB_HD void begin_statement(uint32_t pos) {
    c.A = A_C;
    c.B = B_N;
    c.pk = PK_NONE;
    c.escaped = 0;
    c.esc = 0;
    c.prev_esc = 0;
    c.prev = 0;
    c.stmt_start = pos;
    c.tok_start = pos;
    c.nseg_seen++;
    if (!counting) sink.seg_begin(c.seg, c.count, c.stmt_start);
  }

  // Process the byte at absolute offset pos.  `first_of_batch` suppresses the EOF flush of a
  // non-existent previous statement.
  NUTDB_HD void step(uint32_t pos, uint8_t b, bool is_boundary, bool first_of_batch) {
    if (is_boundary) {
      if (!first_of_batch) {
        flush_eof(pos);
        c.seg++;
      }
      begin_statement(pos);
    }
    const uint8_t ev = a_event(b, c.prev, c.esc != 0);
    const uint8_t a0 = c.A;
    const uint8_t a1 = a_next(a0, ev);
    const bool lc_end = (a0 == A_LC && ev == EV_NL);
    const bool incode = a0 <= A_CX || lc_end;
    uint8_t cl = CL_RESET;

    if (incode) {
      cl = b_class(T, b, c.prev);
      const uint8_t pr = T.prop[b];
      if (lc_end) {
        emit_comment_lc(pos);
        c.pk = PK_NONE;
      }
      // ---- does the token in progress continue with this byte? ----
      bool cont = false;
      switch (c.pk) {
        case PK_NONE: break;
        case PK_WS: cont = (pr & PR_WS) != 0; break;
        case PK_IDENT:
        case PK_CFG: cont = (pr & PR_WORD) != 0; break;
        case PK_QP:
        case PK_NUM:
          switch (c.B) {
            case B_Z: cont = (pr & PR_DIGIT) || cl == CL_XL || cl == CL_DOT; break;
            case B_I: cont = (pr & PR_DIGIT) || cl == CL_DOT; break;
            case B_D0:
            case B_F: cont = (pr & PR_DIGIT) != 0; break;
            case B_H: cont = (pr & PR_HEX) != 0; break;
            default: cont = false;
          }
          break;
        case PK_OP:
          if (cl == CL_LT_P || cl == CL_GT_P || cl == CL_EQ_P) {
            // two-char operator completes here (mod.rs:393-428)
            uint8_t p = c.prev;
            uint8_t tt = p == '!' ? NUTDB_TT_NotEq
                         : p == '<' ? (b == '=' ? NUTDB_TT_LtEq : b == '>' ? NUTDB_TT_NotEq : NUTDB_TT_BitLShift)
                                    : (b == '=' ? NUTDB_TT_GtEq : NUTDB_TT_BitRShift);
            emit(tt, pos - 1, pos + 1);
            c.pk = PK_NONE;
            cont = true;
          }
          break;
        case PK_DASH:
          if (ev == EV_DD && a0 == A_C) {  // "--": line comment starts at the first dash
            c.pk = PK_LC;
            c.tok_start = pos - 1;
            cont = true;
          }
          break;
        case PK_SLASH:
          if (ev == EV_SLST && a0 == A_C) {  // "/*"
            c.pk = PK_BC;
            c.tok_start = pos - 1;
            cont = true;
          }
          break;
        case PK_SQ:
        case PK_DQ:
          // previous byte was a quote that left string context.  A quote of the same kind right
          // after it is the second half of '' / "" (mod.rs:134-137); anything else ends the literal.
          if ((c.pk == PK_SQ && ev == EV_SQ) || (c.pk == PK_DQ && ev == EV_DQ)) {
            c.escaped = 1;
            cont = true;
          }
          break;
        default: break;
      }
      if (!cont) {
        finish_code_token(pos, (pr & PR_IDENT_END) != 0, (pr & PR_NUM_END) != 0);
        // ---- the byte starts fresh ----
        switch (cl) {
          case CL_WS:
            c.pk = PK_WS;
            c.tok_start = pos;
            break;
          case CL_PUNCT:
            if (ev == EV_SQ) {
              c.pk = PK_SQ;
              c.tok_start = pos;
              c.escaped = 0;
            } else if (ev == EV_DQ) {
              c.pk = PK_DQ;
              c.tok_start = pos;
              c.escaped = 0;
            } else if (ev == EV_BT) {
              c.pk = PK_BT;
              c.tok_start = pos;
            } else if (b == '-') {
              c.pk = PK_DASH;
              c.tok_start = pos;
            } else if (b == '/') {
              c.pk = PK_SLASH;
              c.tok_start = pos;
            } else {
              uint8_t tt = T.single_tt[b];
              if (tt != 0xFF) emit(tt, pos, pos + 1);
              else poison(pos, NUTDB_LE_INVALID_CHAR);  // an escaped quote in code context (after a '\\' poison)
            }
            break;
          case CL_LT: case CL_GT: case CL_BANG: case CL_LT_P: case CL_GT_P:
            c.pk = PK_OP;
            c.tok_start = pos;
            break;
          case CL_EQ: case CL_EQ_P: emit(NUTDB_TT_Eq, pos, pos + 1); break;
          case CL_ZERO: case CL_DIG: case CL_DOT: case CL_DOL:
            c.pk = PK_NUM;
            c.tok_start = pos;
            break;
          case CL_HEXL: case CL_XL: case CL_LET: case CL_AT:
            c.pk = PK_IDENT;
            c.tok_start = pos;
            break;
          default:  // CL_BAD
            poison(pos, NUTDB_LE_INVALID_CHAR);
            break;
        }
      }
    } else {
      // ---- inside a string / quoted identifier / comment ----
      switch (a0) {
        case A_SQ:
        case A_DQ:
          if (b == '\\') {
            c.escaped = 1;
          } else if (b == '\r') {
            if (!c.esc) poison(pos, NUTDB_LE_STR_CR);
          } else if (b == '\n') {
            if (!c.esc && !(c.prev == '\r' && c.prev_esc)) poison(pos, NUTDB_LE_STR_LF);
          }
          break;
        case A_BT:
          if (ev == EV_BT) {
            if (pos == c.tok_start + 1) poison(pos, NUTDB_LE_BT_EMPTY);  // mod.rs:323-329
            else emit(NUTDB_TT_DelimitedIdentifier, c.tok_start + 1, pos);
            c.pk = PK_NONE;
          } else if (ev == EV_NL) {
            poison(pos, pos == c.tok_start + 1 ? NUTDB_LE_BT_EMPTY : NUTDB_LE_BT_NL);
          }
          break;
        case A_BC:
          if (ev == EV_STSL) {
            emit_comment_bc(pos);
            c.pk = PK_NONE;
          }
          break;
        default: break;  // A_LC body, A_BC0
      }
    }
    c.B = b_next(c.B, cl);
    c.A = a1;
    c.prev_esc = c.esc;
    c.esc = (b == '\\' && !c.esc) ? 1 : 0;
    c.prev = b;
  }
};

// ------------------------------------------------------------------------------------------
// What phases A and B of the kernel compute per chunk (shared with the host emulation).
// ------------------------------------------------------------------------------------------
// backslash parity in front of `pos`: walks back over the run, never across a statement start
template <class Src>
NUTDB_HD void entry_lookback(Src& src, uint32_t pos, uint32_t batch_begin, uint8_t& prev, uint8_t& esc,
                             uint8_t& prev_esc) {
  prev = 0;
  esc = 0;
  prev_esc = 0;
  if (pos == batch_begin) return;
  prev = src.byte(pos - 1);
  // a statement starts at pos: the walker still needs the last byte of the PREVIOUS statement to
  // finish its final token in flush_eof(); esc/prev_esc are reset by begin_statement() anyway
  if (src.boundary(pos)) return;
  uint32_t n = 0;  // backslashes directly before pos
  uint32_t p = pos;
  while (p > batch_begin && !src.boundary(p) && src.byte(p - 1) == '\\') {
    n++;
    p--;
  }
  esc = (uint8_t)(n & 1u);
  // prev_esc = parity of the run in front of pos-1
  if (prev == '\\') {
    prev_esc = (uint8_t)((n - 1) & 1u);
  } else {
    uint32_t m = 0;
    p = pos - 1;
    while (p > batch_begin && !src.boundary(p) && src.byte(p - 1) == '\\') {
      m++;
      p--;
    }
    prev_esc = (uint8_t)(m & 1u);
  }
}


// pk at a chunk entry is re-derived from (A, B, previous byte): identical to the walker's own pk in
// every live region (before the first poison of a statement).
NUTDB_HD uint8_t derive_pk(const LexTables& T, uint8_t A, uint8_t B, uint8_t prev) {
  switch (A) {
    case A_SQ: return PK_SQ;
    case A_DQ: return PK_DQ;
    case A_BT: return PK_BT;
    case A_LC: return PK_LC;
    case A_BC0:
    case A_BC: return PK_BC;
    case A_CX: return PK_NONE;
    default: break;
  }
  switch (B) {
    case B_W: return PK_IDENT;  // or PK_CFG: told apart at finish by the first byte
    case B_Z: case B_I: case B_D0: case B_F: case B_H: return PK_NUM;  // or PK_QP, same
    case B_OP: return PK_OP;
    default: break;
  }
  if (prev == 0) return PK_NONE;
  if (T.prop[prev] & PR_WS) return PK_WS;
  if (prev == '-') return PK_DASH;
  if (prev == '/') return PK_SLASH;
  if (prev == '\'') return PK_SQ;
  if (prev == '"') return PK_DQ;
  return PK_NONE;
}

// ---- phase A: transition function of automaton A over [begin, end) ----
template <class Src>
NUTDB_HD uint32_t chunk_sim_A(const LexTables& T, Src& src, uint32_t begin, uint32_t end) {
  uint8_t prev, esc, prev_esc;
  entry_lookback(src, begin, 0u, prev, esc, prev_esc);
  uint32_t run = NUTDB_VEC8_ID;
  for (uint32_t pos = begin; pos < end; pos++) {
    uint8_t b = src.byte(pos);
    if (src.boundary(pos)) {
      run = 0u;  // constant function -> A_C
      prev = 0;
      esc = 0;
    }
    uint8_t ev = a_event(b, prev, esc != 0);
    run = vec8_then_row(run, T.a_row[ev][0], T.a_row[ev][1]);
    esc = (b == '\\' && !esc) ? 1 : 0;
    prev = b;
  }
  return run;
}

// ---- phase B: transition function of automaton B over [begin, end), A entry state known ----
template <class Src>
NUTDB_HD uint32_t chunk_sim_B(const LexTables& T, Src& src, uint32_t begin, uint32_t end, uint8_t entryA) {
  uint8_t prev, esc, prev_esc;
  entry_lookback(src, begin, 0u, prev, esc, prev_esc);
  uint32_t run = NUTDB_VEC8_ID;
  uint8_t A = entryA;
  for (uint32_t pos = begin; pos < end; pos++) {
    uint8_t b = src.byte(pos);
    if (src.boundary(pos)) {
      run = 0u;  // constant function -> B_N
      A = A_C;
      prev = 0;
      esc = 0;
    }
    uint8_t ev = a_event(b, prev, esc != 0);
    bool incode = A <= A_CX || (A == A_LC && ev == EV_NL);
    uint8_t cl = incode ? b_class(T, b, prev) : (uint8_t)CL_RESET;
    run = vec8_then_row(run, T.b_row[cl][0], T.b_row[cl][1]);
    A = a_next(A, ev);
    esc = (b == '\\' && !esc) ? 1 : 0;
    prev = b;
  }
  return run;
}

// ---- phase C: what a chunk contributes to the running (count, statement, token-in-progress) ----
struct CSum {
  uint32_t count;       // tokens emitted
  uint32_t nseg;        // statement starts seen
  uint32_t tok_start;   // valid if has_tok
  uint32_t stmt_start;  // valid if nseg > 0
  uint8_t has_tok;      // a token (or statement) started in the chunk
  uint8_t escaped;      // escaped flag of the token in progress at the end (since has_tok, or OR-ed in)
};
NUTDB_HD CSum csum_identity() { return CSum{0u, 0u, 0u, 0u, 0, 0}; }
// first a, then b
NUTDB_HD CSum csum_then(const CSum& a, const CSum& b) {
  CSum r;
  r.count = a.count + b.count;
  r.nseg = a.nseg + b.nseg;
  r.has_tok = a.has_tok | b.has_tok;
  r.tok_start = b.has_tok ? b.tok_start : a.tok_start;
  r.escaped = b.has_tok ? b.escaped : (uint8_t)(a.escaped | b.escaped);
  r.stmt_start = b.nseg ? b.stmt_start : a.stmt_start;
  return r;
}

struct CountSink {
  NUTDB_HD void token(uint32_t, uint8_t, uint32_t, uint32_t, uint8_t) {}
  NUTDB_HD void seg_begin(uint32_t, uint32_t, uint32_t) {}
  NUTDB_HD void seg_end(uint32_t, uint32_t, uint32_t) {}
};

#define NUTDB_NO_TOK 0xFFFFFFFFu

// Runs the walker over [begin, end) (and the final EOF flush if the batch ends inside or at the
// end of this chunk).  carry.count / seg / tok_start / stmt_start / escaped come from the scan.
template <bool EmitAll, class Src, class Sink>
NUTDB_HD LexCarry chunk_walk(const LexTables& T, Src& src, Sink& sink, uint32_t begin, uint32_t end, uint32_t batch_end,
                             uint8_t entryA, uint8_t entryB, const CSum& prefix, bool counting) {
  LexCarry c;
  c.A = entryA;
  c.B = entryB;
  entry_lookback(src, begin, 0u, c.prev, c.esc, c.prev_esc);
  c.pk = derive_pk(T, entryA, entryB, c.prev);
  c.count = prefix.count;
  c.seg = prefix.nseg ? prefix.nseg - 1 : 0;
  c.stmt_start = prefix.stmt_start;
  c.tok_start = counting ? NUTDB_NO_TOK : prefix.tok_start;
  c.escaped = counting ? 0 : prefix.escaped;
  Walker<EmitAll, Src, Sink> w(T, src, sink, c);
  w.counting = counting;
  uint32_t stop = end < batch_end ? end : batch_end;
  for (uint32_t pos = begin; pos < stop; pos++) w.step(pos, src.byte(pos), src.boundary(pos), pos == 0);
  if (batch_end > begin && batch_end <= end) w.flush_eof(batch_end);
  return w.c;
}

template <bool EmitAll, class Src>
NUTDB_HD CSum chunk_count_t(const LexTables& T, Src& src, uint32_t begin, uint32_t end, uint32_t batch_end,
                            uint8_t entryA, uint8_t entryB) {
  CountSink sink;
  CSum zero = csum_identity();
  LexCarry c = chunk_walk<EmitAll>(T, src, sink, begin, end, batch_end, entryA, entryB, zero, true);
  CSum s;
  s.count = c.count;
  s.nseg = c.nseg_seen;
  s.has_tok = c.tok_start != NUTDB_NO_TOK;
  s.tok_start = c.tok_start;
  s.escaped = c.escaped;
  s.stmt_start = c.stmt_start;
  return s;
}
template <class Src>
NUTDB_HD CSum chunk_count(const LexTables& T, Src& src, uint32_t begin, uint32_t end, uint32_t batch_end,
                          uint8_t entryA, uint8_t entryB, bool emit_all) {
  return emit_all ? chunk_count_t<true>(T, src, begin, end, batch_end, entryA, entryB)
                  : chunk_count_t<false>(T, src, begin, end, batch_end, entryA, entryB);
}

}  // namespace nlex
