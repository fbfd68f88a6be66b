// lex_tables.hpp -- host-side construction of nlex::LexTables (byte classes, the two automata as
// per-input transition rows, and the keyword perfect hash).  The rows are DERIVED from a_next /
// b_next in lex_core.cuh, so the scanned transition functions and the walker cannot disagree.
#pragma once
#include <cstring>
#include <stdexcept>

#include "lex_core.cuh"

namespace nlex {

// keyword.rs:13-148, declaration order (id = index + 1)
static const char* const KEYWORD_TEXT[NUTDB_KW_COUNT] = {
    "by", "as", "on", "from", "intersect", "union", "all", "except", "distinct", "with", "select", "join", "where",
    "group", "having", "order", "limit", "offset", "using", "ties", "asc", "desc", "explain", "insert", "into",
    "values", "create", "primary", "key", "comment", "update", "default", "check", "describe", "drop", "alter", "add",
    "rename", "first", "after", "truncate", "optimize", "set", "database", "table", "view", "column", "index",
    "constraint", "partition", "null", "true", "false", "and", "or", "xor", "not", "in", "exists", "if", "case",
    "when", "then", "else", "end", "is", "between", "like", "ilike", "interval", "second", "minute", "hour", "day",
    "month", "year", "int8", "int16", "int32", "int64", "int128", "uint8", "uint16", "uint32", "uint64", "uint128",
    "serial32", "serial64", "serial128", "userial32", "userial64", "userial128", "decimal32", "decimal64", "float32",
    "float64", "boolean", "chars", "string", "uuid", "date", "datetime", "array", "enum", "tuple", "map",
    "dictionary", "nullable", "inner", "outer", "left", "right", "full", "semi", "anti"};

inline void build_lex_tables(LexTables& T) {
  std::memset(&T, 0, sizeof(T));
  for (int i = 0; i < 256; i++) {
    uint8_t c = (uint8_t)i;
    uint8_t cl = CL_BAD, pr = 0, tt = 0xFF;
    bool lower = c >= 'a' && c <= 'z', upper = c >= 'A' && c <= 'Z', digit = c >= '0' && c <= '9';
    if (c == ' ' || c == '\t' || c == '\n' || c == '\r') {
      cl = CL_WS;
      pr |= PR_WS;
    } else if (digit) {
      cl = c == '0' ? CL_ZERO : CL_DIG;
      pr |= PR_WORD | PR_DIGIT | PR_HEX;
    } else if (lower || upper || c == '_') {
      uint8_t l = (uint8_t)(c | 0x20);
      bool hex = (lower || upper) && l >= 'a' && l <= 'f';
      cl = hex ? CL_HEXL : ((c == 'x' || c == 'X') ? CL_XL : CL_LET);
      pr |= PR_WORD;
      if (hex) pr |= PR_HEX;
    } else {
      switch (c) {
        case '(': tt = NUTDB_TT_LParen; break;
        case ')': tt = NUTDB_TT_RParen; break;
        case '[': tt = NUTDB_TT_LBracket; break;
        case ']': tt = NUTDB_TT_RBracket; break;
        case '{': tt = NUTDB_TT_LBrace; break;
        case '}': tt = NUTDB_TT_RBrace; break;
        case ',': tt = NUTDB_TT_Comma; break;
        case ':': tt = NUTDB_TT_Colon; break;
        case '+': tt = NUTDB_TT_Plus; break;
        case '*': tt = NUTDB_TT_Mul; break;
        case '%': tt = NUTDB_TT_Mod; break;
        case '&': tt = NUTDB_TT_BitAnd; break;
        case '|': tt = NUTDB_TT_BitOr; break;
        case '^': tt = NUTDB_TT_BitXor; break;
        case '~': tt = NUTDB_TT_BitNot; break;
        case ';': tt = NUTDB_TT_SemiColon; break;
        default: break;
      }
      if (tt != 0xFF || c == '-' || c == '/' || c == '\'' || c == '"' || c == '`') cl = CL_PUNCT;
      else if (c == '<') cl = CL_LT;
      else if (c == '>') cl = CL_GT;
      else if (c == '!') cl = CL_BANG;
      else if (c == '=') cl = CL_EQ;
      else if (c == '.') cl = CL_DOT;
      else if (c == '@') cl = CL_AT;
      else if (c == '$') cl = CL_DOL;
    }
    // tokenizer/mod.rs:486-503 (identifier) and :506-543 (query parameter == numeric)
    const char* ident_end = "+-*/%&|^><=!.,;[](){}\t\n\r ";
    const char* num_end = "+-*/%&|^><=!,:;])}\t\n\r ";
    if (c != 0 && std::strchr(ident_end, c)) pr |= PR_IDENT_END;
    if (c != 0 && std::strchr(num_end, c)) pr |= PR_NUM_END;
    T.base_cls[i] = cl;
    T.prop[i] = pr;
    T.single_tt[i] = tt;
    T.tt0[i] = (pr & PR_DIGIT) ? (uint8_t)NUTDB_TT_IntegerLiteral
               : (pr & PR_WORD) ? (uint8_t)NUTDB_TT_KeywordOrIdentifier
               : c == '.' ? (uint8_t)NUTDB_TT_Dot
               : c == '-' ? (uint8_t)NUTDB_TT_Minus
               : c == '/' ? (uint8_t)NUTDB_TT_Div
               : c == '<' ? (uint8_t)NUTDB_TT_Lt
               : c == '>' ? (uint8_t)NUTDB_TT_Gt
               : c == '=' ? (uint8_t)NUTDB_TT_Eq : tt;
  }
  for (int ev = 0; ev < EV_COUNT; ev++) {
    uint32_t lo = 0, hi = 0;
    for (int s = 0; s < 4; s++) lo |= (uint32_t)a_next((uint8_t)s, (uint8_t)ev) << (8 * s);
    for (int s = 4; s < 8; s++) hi |= (uint32_t)a_next((uint8_t)s, (uint8_t)ev) << (8 * (s - 4));
    T.a_row[ev][0] = lo;
    T.a_row[ev][1] = hi;
  }
  for (int cl = 0; cl < CL_COUNT; cl++) {
    uint32_t lo = 0, hi = 0;
    for (int s = 0; s < 4; s++) lo |= (uint32_t)b_next((uint8_t)s, (uint8_t)cl) << (8 * s);
    for (int s = 4; s < 8; s++) hi |= (uint32_t)b_next((uint8_t)s, (uint8_t)cl) << (8 * (s - 4));
    T.b_row[cl][0] = lo;
    T.b_row[cl][1] = hi;
  }
  // keywords + perfect hash: slot = (c0*m0 + c1*m1 + c[n-1]*m2 + c[n-2]*m3 + n) mod 512, all 115 words distinct
  for (int k = 0; k < NUTDB_KW_COUNT; k++) {
    size_t n = std::strlen(KEYWORD_TEXT[k]);
    for (size_t i = 0; i < n; i++) {  // keyword_lookup_words folds case with `| 0x20`: exact for [a-z0-9] only
      const char c = KEYWORD_TEXT[k][i];
      if (!((c >= 'a' && c <= 'z') || (c >= '0' && c <= '9')) || n < 2 || n > 10)
        throw std::runtime_error("keyword outside [a-z0-9]{2,10}");
    }
    T.kw_len[k + 1] = (uint8_t)n;
    std::memcpy(T.kw_text[k + 1], KEYWORD_TEXT[k], n);
  }
  bool found = false;
  for (uint32_t m3 = 1; m3 < 32 && !found; m3++)
    for (uint32_t m0 = 1; m0 < 64 && !found; m0++)
      for (uint32_t m1 = 1; m1 < 64 && !found; m1++)
        for (uint32_t m2 = 1; m2 < 64 && !found; m2++) {
          uint8_t slot[NUTDB_KW_SLOTS];
          std::memset(slot, 0, sizeof(slot));
          bool ok = true;
          for (int k = 0; k < NUTDB_KW_COUNT && ok; k++) {
            const char* w = KEYWORD_TEXT[k];
            uint32_t n = (uint32_t)std::strlen(w);
            uint32_t h = ((uint8_t)w[0] * m0 + (uint8_t)w[1] * m1 + (uint8_t)w[n - 1] * m2 + (uint8_t)w[n - 2] * m3 + n) &
                         (NUTDB_KW_SLOTS - 1);
            if (slot[h]) ok = false;
            else slot[h] = (uint8_t)(k + 1);
          }
          if (ok) {
            std::memcpy(T.kw_slot, slot, sizeof(slot));
            T.kw_mul[0] = m0;
            T.kw_mul[1] = m1;
            T.kw_mul[2] = m2;
            T.kw_mul[3] = m3;
            found = true;
          }
        }
  if (!found) throw std::runtime_error("nutdb: keyword perfect hash not found");
}

}  // namespace nlex
