// parse_fast.cuh -- table-driven parser for the statement shapes that dominate query logs.
//
// The bytecode automaton in parse_core.cuh is exact for the whole grammar but pays an
// interpreter's price (fetch/dispatch/stack per grammar step).  This file parses the COMMON shapes:
//
//   SELECT [DISTINCT] items [FROM name [AS a] {[INNER|LEFT|RIGHT|FULL ..] JOIN name [AS a] ON e | USING (names)}] [WHERE e] [GROUP BY items] [HAVING e] [ORDER BY item [DESC],..]
//          [LIMIT n [, m | OFFSET m] [WITH TIES]]
//   INSERT INTO name [(names)] VALUES (exprs) {, (exprs)}
//   CREATE TABLE [IF NOT EXISTS] name (name type [DEFAULT e | COMMENT s].. ,..)
//          [PRIMARY KEY es | ORDER BY es | PARTITION BY e | COMMENT s]..
//
// with expressions over identifiers, literals, the binary operators, AND/OR/XOR, [NOT] IN/LIKE/ILIKE,
// [NOT] BETWEEN .. AND .., IS [NOT] NULL, prefix NOT, CASE .. END, parenthesised groups / tuples and function calls
// (operator-precedence parsing with an explicit
// operator stack -- the iterative form of must_parse_expr_tdop, reference mod.rs:1209-1220: an
// operator is reduced when one of equal or lower power arrives, so every operator is
// left-associative exactly as in the reference).
//
// Why a table: one thread parses one statement, so the 32 lanes of a warp sit at 32 different
// places of the grammar.  Straight-line code for every clause gave each lane its own code path:
// ~6 of 32 lanes active per instruction and -- worse -- a 125 KB kernel whose hot half did not fit
// the instruction caches (61 % of the warp stall samples were instruction fetch).  Here the
// grammar outside expressions is DATA: a (state, token class) -> transition record table in shared
// memory (parse_fast_tables.hpp builds it), and one small loop that every lane executes once per token:
//
//     operator phase (only in state X_OPER) -> table look-up -> one of 8 short actions
//
// The whole parser is a few hundred instructions, resident in the instruction cache.
//
// Two instantiations share this code and the table.  The NARROW one (WIDE = false) is the first pass over every
// statement: operator stack of FAST_STACK_DEPTH entries in shared memory, the subset above.  The WIDE one is the second
// pass over what the first declined: its stack lives at the top end of the statement's own compact-node range (nodes
// grow up from the bottom, the stack down from the top; any nesting depth the token count allows), and it also takes
// array literals [..], map literals {k: v, ..}, index access x[i], prefix ~, IF .. THEN .. ELSE .. END and parenthesised
// subqueries (select ..) -- the constructs of the reference's corpus and of deeply nested statements.
//
// It is ALL-OR-NOTHING: on anything outside that subset -- any error, any construct that needs
// constant folding (simplify.rs), literal validation beyond a length check, joins, set
// operations, subqueries, IF/INTERVAL, NOT EXISTS, arrays, maps ... -- try_parse() returns false
// without side effects the caller keeps, and the statement is parsed from scratch by the exact
// automaton.  For a statement it accepts, it emits precisely the nodes the automaton would.
#pragma once
#include "parse_core.cuh"

#ifndef NUTDB_WIDE_LITE
#define NUTDB_WIDE_LITE 0   // experiment switch: 1 compiles the set operations / Enum out of the wide pass
#endif
#ifndef FAST_STACK_DEPTH
#define FAST_STACK_DEPTH 8
#endif

namespace npar {

struct FastStackEntry {
  uint32_t x, y;
};

// ---- token classes: everything the grammar states distinguish ----
enum FastClass : uint8_t {
  FC_OTHER = 0, FC_EOF, FC_SEMI, FC_COMMA, FC_LPAREN, FC_RPAREN, FC_MUL, FC_PLUS, FC_MINUS, FC_BINOP, FC_LBRACKET,
  FC_LBRACE, FC_BITNOT,
  FC_INT, FC_HEX, FC_FLOAT, FC_RAWSTR, FC_ESQ, FC_EDQ, FC_DELIM, FC_DOT,
  // words (token type KeywordOrIdentifier), by keyword id: every class from FC_WORD on is a word class
  FC_WORD, FC_TRUE, FC_FALSE, FC_NULL, FC_NOT, FC_IF, FC_BADPFX, FC_KWBINOP, FC_ISBETWEEN,
  FC_FROM, FC_WHERE, FC_GROUP, FC_BY, FC_HAVING, FC_ORDER, FC_LIMIT, FC_OFFSET, FC_WITH, FC_TIES, FC_AS, FC_DESC,
  FC_INTO, FC_VALUES, FC_TABLE, FC_EXISTS, FC_DEFAULT, FC_COMMENT, FC_PRIMARY, FC_KEY, FC_PARTITION, FC_DISTINCT,
  FC_SETOP, FC_JOIN, FC_INDEX, FC_CONSTRAINT, FC_CHECK, FC_VIEW, FC_UPDATE, FC_SELECT, FC_DTYPE,
  FC_ON, FC_USING, FC_INNER, FC_FULL, FC_LEFT, FC_RIGHT, FC_OUTER, FC_KSEMI, FC_KANTI, FC_CASE,
  FC_COUNT
};
static const uint32_t FC_FIRST_WORD = FC_WORD;

// where the expression being parsed sits in its statement; the state after an expression is FS_AFTER + context
enum FastCtx : uint8_t { C_SEL_ITEM = 0, C_WHERE, C_GROUP_ITEM, C_HAVING, C_ORDER_ITEM, C_INS_VALUE, C_COL_DEFAULT,
                         C_TBL_PK_ITEM, C_TBL_ORDER_ITEM, C_TBL_PART, C_JOIN_ON, C_IDX_EXPR, C_CON_EXPR,
                         C_V_PK_ITEM, C_V_ORDER_ITEM, C_V_PART, C_COUNT };

enum FastState : uint8_t {
  FS_X_OPND = 0, FS_X_OPER, FS_AFTER, FS_AFTER_END = FS_AFTER + C_COUNT - 1,
  FS_SEL0, FS_SEL_D, FS_SEL_ALIAS, FS_SEL_ITEM2, FS_FROM, FS_SRC, FS_SRC2, FS_SRC_Q, FS_SRC2B, FS_SRC_ALIAS, FS_SRC3,
  FS_J_KW, FS_J_OUTER, FS_J_LEFT, FS_J_RIGHT, FS_J_SRC, FS_J_SRC2, FS_J_SRC_Q, FS_J_SRC2B, FS_J_ALIAS, FS_J_ONUSING, FS_J_U_LP, FS_J_U_ID, FS_J_U_SEP,
  FS_CL1, FS_CL2, FS_CL3, FS_CL4, FS_CL5, FS_GROUP_BY, FS_ORDER_BY, FS_GRP_ALIAS, FS_GRP_ITEM2,
  FS_ORD_ALIAS, FS_ORD_ITEM2, FS_ORD_ITEM3,
  FS_LIM1, FS_LIM2, FS_LIM3A, FS_LIM4A, FS_LIM3B, FS_LIM4B, FS_TIES0, FS_TIES1, FS_TIES2, FS_BODY, FS_END_SEL,
  FS_INS0, FS_INS_NAME, FS_INS_AFTER_NAME, FS_INS_COL, FS_INS_COL_SEP, FS_INS_VALUES, FS_INS_ROW0, FS_INS_AFTER_ROW,
  FS_INS_ROWN, FS_INS_END,
  FS_CRE0, FS_CRE1, FS_CRE_IF1, FS_CRE_IF2, FS_CRE_NAME, FS_CRE_LP, FS_COL_BEGIN, FS_DT, FS_DT_END, FS_COL_ATTRS,
  FS_COL_COMMENT, FS_COL_SEP, FS_TBL_ATTRS, FS_TBL_KEY, FS_TBL_ORDER_BY, FS_TBL_PART_BY, FS_TBL_COMMENT, FS_CRE_END,
  FS_WITH0, FS_WITH_AS, FS_WITH_LP, FS_WITH_SEP, FS_IDX_NAME, FS_CON_NAME, FS_CON_CHECK,
  FS_CRV1, FS_CRV_IF1, FS_CRV_IF2, FS_CRV_NAME, FS_VIEW_ATTRS, FS_V_UPD, FS_V_STRAT, FS_V_KEY, FS_V_ORDER_BY, FS_V_PART_BY,
  FS_V_COMMENT,
  FS_FINAL,
  FS_COUNT
};

enum FastAct : uint32_t { FA_BAIL = 0, FA_STEP, FA_IDENT, FA_NEG, FA_OPEN, FA_DTYPE, FA_DTEND, FA_ROWEND, FA_ACCEPT, FA_PUSH, FA_CASE,
                          // the wide instantiation only (the narrow one declines them):
                          FA_SETOP, FA_INTERVAL, FA_SRC_SUBQ, FA_SUBQ_END };

// ---- transition record: two words ----
// lo: act[0:4) next[4:11) adv[11] emit[12:15) kind[15:23) sub[23:27) auxbit[27] auxreg[28] subreg[29]
// hi: pre[0:2) post[2:4) setcur[4] check[5:8) look[8:10) setctx[10] ctx[11:15) bit[15:23) clr[23] inccnt[24] setaux[25]
//     setjr[26] jr[27:30)   (jr = the join type, kept in a register until the JOIN node is emitted: subreg)
//     popnode[30]           (the node emitted last is withdrawn: the qualifier of `db.table`)
enum : uint32_t { FE_NONE = 0, FE_LEAF_TOK, FE_LEAF_NOTOK, FE_NODE_M0, FE_NODE_M1, FE_NODE_ZERO };
enum : uint32_t { FK_NONE = 0, FK_INT_W0, FK_INT_W1, FK_INT_W2, FK_STR, FK_FNCALL };
enum : uint32_t { FL_NONE = 0, FL_NODOT, FL_NODOT_NOLP, FL_NOLP };
static const uint32_t FAST_MAX_REC = 192;
static const uint32_t FAST_HI_UNCOMMON = 0x7FFFFFE0u;  // every hi field except pre / post / setcur

// Token-indexed tables take ONE index for both kinds of token: the token type, or 64 + keyword id for a word.
struct FastTables {
  uint16_t op[192];     // power | op << 4 | bail << 12   (token_power, mod.rs:1895-1927)
  uint8_t cls[192];     // -> FastClass
  uint8_t trans[FS_COUNT][FC_COUNT];  // -> record index (0 = bail)
  uint32_t rec_lo[FAST_MAX_REC], rec_hi[FAST_MAX_REC];
  uint32_t rec_hdr[FAST_MAX_REC];  // first word of the compact node a record emits: kind | sub << 8 | aux bit << 16
};
NUTDB_HD uint32_t fast_token_index(uint32_t ty, uint32_t kw) { return ty == NUTDB_TT_KeywordOrIdentifier ? 64u + kw : ty; }

template <class Tok, class Nodes, class Text, bool WIDE = false>
struct FastParser {
  const FastTables* F;
  Tok tok;
  Nodes nd;
  Text text;
  FastStackEntry* stk;  // the caller's memory: entry i of this thread at stk[i * stride] (narrow: shared memory, strided
  int32_t stride;       // over the CTA's threads; wide: stride -1 from the last slot of the statement's node range)
  // operator / bracket stack entries: x = type | power << 4 | op << 8 | left kind << 14 | item count << 22
  // (E_DT: x = type | compound sub << 4), y = start of the left operand (operators) / node count at the opening.
  // The power of an entry is the min_power its right operand is parsed with (must_parse_expr_tdop, mod.rs:1209):
  // an arriving operator of equal or lower power completes it.  E_NOT / E_BITNOT = prefix NOT / ~, E_BTW1 / E_BTW2 =
  // [NOT] BETWEEN before / after its AND (op = FnName Between 3 / NotBetween 4).  Everything from E_PAREN on is a
  // BRACKET: power 0, closed by its own terminator.
  // E_CASE = CASE [scrutinee] WHEN .. THEN .. [ELSE ..] END: x = type | position << 8 | FnName << 10 (position: 0 after
  // the scrutinee, 1 after a condition, 2 after a result, 3 after the ELSE expression).
  // Wide only: E_IF (x = type | position << 8: 0 after the condition, 1 after THEN's value, 2 after ELSE's),
  // E_BRACKET = [items], E_MAP = {k: v, ..} (position << 8: 0 after a key, 1 after a value), E_INDEX = left[ (y = start of
  // left), E_SUBQ = (select ..: x = type | ctx << 8 | join register << 12 | SUBQ_CALL / SUBQ_SOURCE, y = the outer query's
  // base; the entry below it holds the outer m0 / m1), E_UNION = a set operation waiting for its right query (x = type |
  // UnionTypePower << 4 | UnionType << 8, y = first node of its left query).
  enum : uint32_t { E_OP = 0, E_NOT = 1, E_BITNOT = 2, E_BTW1 = 3, E_BTW2 = 4, E_PAREN = 5, E_CALL = 6, E_DT = 7, E_CASE = 8,
                    E_IF = 9, E_BRACKET = 10, E_MAP = 11, E_INDEX = 12, E_SUBQ = 13, E_UNION = 14 };
  enum : uint32_t { SPEC_NONE = 0, SPEC_BAIL = 1, SPEC_NOT = 2, SPEC_IS = 3, SPEC_BETWEEN = 4 };  // FastTables::op >> 12
  enum : uint32_t { X_POWER = 4, X_OP = 8, X_LKIND = 14, X_COUNT = 22, X_COUNT_MAX = 1023 };
  enum : uint32_t { SUBQ_CALL = 1u << 15, SUBQ_SOURCE = 1u << 16, SUBQ_CTE = 1u << 17, SUBQ_VIEW = 1u << 18 };  // E_SUBQ: what the subquery is part of
  static const uint32_t DEPTH = WIDE ? 0x7FFFFFFFu : (uint32_t)FAST_STACK_DEPTH;  // (wide: bounded by the node range, see try_parse)

  NUTDB_HD FastParser(const FastTables* ft, const Tok& tk, const Nodes& nodes, const Text& tx, FastStackEntry* stack,
                      int32_t stack_stride)
      : F(ft), tok(tk), nd(nodes), text(tx), stk(stack), stride(stack_stride) {}

  NUTDB_HD static bool is_literal(uint32_t kind) { return kind >= NUTDB_NK_LIT_INT && kind <= NUTDB_NK_LIT_INTERVAL; }
  NUTDB_HD static bool ident_string(uint32_t y) {  // must_parse_identifier_string (mod.rs:1682)
    return y == NUTDB_TT_KeywordOrIdentifier || y == NUTDB_TT_DelimitedIdentifier;
  }
  // integer_from_str! cannot fail for 1..safe digits; the lexer stores min(len, 255) in the kw byte
  NUTDB_HD static bool int_ok(uint32_t y, uint32_t len, uint32_t width_) {
    const bool hex = y == NUTDB_TT_HexLiteral;
    const uint32_t safe = width_ == 0 ? 2u : width_ == 1 ? (hex ? 16u : 19u) : (hex ? 32u : 38u);
    return len >= 1 && len <= safe;
  }
  // Would unescaping this escaped string literal succeed (literal.rs:45-102)?  Only a backslash-u escape can be
  // rejected (InvalidEscapedUnicode), and a lone trailing backslash makes the reference panic (literal.rs:63); both
  // go to the automaton, which reports them.  Byte-wise is exact: every character that matters is ASCII, and the
  // continuation bytes of a skipped multi-byte character are inert.  `quote` doubles as an escape ('' / "").
  NUTDB_HD bool string_ok(uint32_t i, uint32_t quote) {
    uint32_t p = tok.start(i);
    const uint32_t e = tok.end(i);
    while (p < e) {
      p = text.skip_plain(p, e, quote);  // (four bytes at a time over what is neither a backslash nor the quote)
      if (p >= e) break;
      const uint32_t c = text.raw(p++);  // (every index below is < e <= the statement's length)
      if (c == quote) {
        p++;  // chars.next()
        continue;
      }
      if (c != '\\') continue;
      if (p >= e) return false;  // the reference panics here
      if (text.raw(p++) != 'u') continue;
      if (p >= e || text.raw(p++) != '{') continue;  // plain 'u'; the character after it is dropped
      bool ok = true, any = false, first = true;
      uint64_t v = 0;
      while (p < e) {  // take_while(|&ch| ch != '}') then u32::from_str_radix(.., 16) and char::from_u32
        const uint32_t h = text.raw(p++);
        if (h == '}') break;
        if (first && h == '+') {
          first = false;
          continue;
        }
        first = false;
        uint32_t d;
        if (h >= '0' && h <= '9') d = h - '0';
        else if (h >= 'a' && h <= 'f') d = h - 'a' + 10;
        else if (h >= 'A' && h <= 'F') d = h - 'A' + 10;
        else {
          ok = false;
          continue;
        }
        any = true;
        if (v <= 0xFFFFFFFFull) v = v * 16 + d;
      }
      if (!ok || !any || v > 0x10FFFF || (v >= 0xD800 && v <= 0xDFFF)) return false;
    }
    return true;
  }

  // ---- constant folding (simplify.rs), wide pass only: the narrow pass declines whatever would fold ----
  // literal == literal (Literal: PartialEq): 0 / 1, or 2 = not decided here (floats, escaped strings: the automaton)
  NUTDB_HD uint32_t lit_equal_simple(const CNode& x, const CNode& y) {
    if (x.kind != y.kind) return 0u;
    if (x.kind == NUTDB_NK_LIT_NULL) return 1u;
    if (x.kind == NUTDB_NK_LIT_BOOL) return x.sub == y.sub ? 1u : 0u;
    if (x.x == NUTDB_CN_NOTOK || y.x == NUTDB_CN_NOTOK) return 2u;
    const uint32_t xa = tok.start(x.x), xb = tok.end(x.x), ya = tok.start(y.x), yb = tok.end(y.x);
    if (x.kind == NUTDB_NK_LIT_INT) {  // sign + magnitude (at most 38 / 32 digits: int_ok let them in)
      if ((x.sub & 1) != (y.sub & 1)) return 0u;
      u128_t v[2] = {0, 0};
      for (int k = 0; k < 2; k++) {
        const bool hex = ((k ? y.aux : x.aux) & 1) != 0;
        const uint32_t e = k ? yb : xb;
        for (uint32_t p = k ? ya : xa; p < e; p++) {
          const uint8_t ch = text.byte(p);
          const uint32_t d = ch <= '9' ? (uint32_t)(ch - '0') : (uint32_t)((ch | 0x20) - 'a' + 10);
          v[k] = v[k] * (hex ? 16u : 10u) + d;
        }
      }
      return v[0] == v[1] ? 1u : 0u;
    }
    if (x.kind == NUTDB_NK_LIT_STR && x.sub == 0 && y.sub == 0) {  // raw strings: the bytes
      if (xb - xa != yb - ya) return 0u;
      for (uint32_t i = 0; i < xb - xa; i++)
        if (text.byte(xa + i) != text.byte(ya + i)) return 0u;
      return 1u;
    }
    return 2u;
  }
  // removes leaf node i, sliding [i + 1, n) down by one and re-basing the subtree starts stored in them
  NUTDB_HD void remove_leaf(uint32_t i, uint32_t& n) {
    for (uint32_t j = i + 1; j < n; j++) {
      CNode x = nd.get(j);
      if (x.kind >= NUTDB_NK_FIRST_INTERIOR) x.x -= 1;
      nd.set(j - 1, x);
    }
    n -= 1;
  }

  // Enum('a' [= n], ..) from its `Enum` word at token t: the binds as NK_STR [NK_NUM] leaves under one DT_COMPOUND
  // (must_parse_enum_binds, mod.rs:1799-1813).  (Out-of-line helpers measured 50 % SLOWER: the calls spill the parser state.)
  NUTDB_HD bool parse_enum(uint32_t& t, uint32_t& n, uint32_t sp, uint32_t cap) {
    const uint32_t m = n;
    uint32_t u = t + 2u;
    for (;;) {
      if (n + sp + 6u > cap) return false;
      const uint32_t pu = tok.pair_at(u), ps = pu & 255u;
      uint32_t ssub;
      if (ps == NUTDB_TT_RawStringLiteral) ssub = 0;
      else if (ps == NUTDB_TT_EscapedSQStringLiteral) ssub = 1;
      else if (ps == NUTDB_TT_EscapedDQStringLiteral) ssub = 2;
      else return false;
      if (ssub && (pu >> 8) != 1u && !string_ok(u, ssub == 1 ? '\'' : '"')) return false;
      CNode c;
      c.kind = NUTDB_NK_STR;
      c.sub = (uint8_t)ssub;
      c.aux = 0;
      c.x = u;
      nd.set(n++, c);
      u++;
      if ((tok.pair_at(u) & 255u) == NUTDB_TT_Eq) {
        const uint32_t pi = tok.pair_at(u + 1u);
        const uint32_t tyi = pi & 255u;
        if (tyi != NUTDB_TT_IntegerLiteral && tyi != NUTDB_TT_HexLiteral) return false;
        if (!int_ok(tyi, pi >> 8, 1)) return false;
        c.kind = NUTDB_NK_NUM;
        c.sub = 0;
        c.aux = tyi == NUTDB_TT_HexLiteral ? 1 : 0;
        c.x = u + 1u;
        nd.set(n++, c);
        u += 2u;
      }
      if ((tok.pair_at(u) & 255u) != NUTDB_TT_Comma) break;
      u++;
    }
    if ((tok.pair_at(u) & 255u) != NUTDB_TT_RParen) return false;
    CNode c;
    c.kind = NUTDB_NK_DT_COMPOUND;
    c.sub = 1;
    c.aux = 0;
    c.x = m;
    nd.set(n++, c);
    t = u + 1u;
    return true;
  }

  // parse_stmt (mod.rs:128-180).  true: res describes a successful parse with node_count nodes emitted.
  NUTDB_HD bool try_parse(ParseResult& res) {
    uint32_t p = tok.pair_at(0);
    if ((p & 255u) != NUTDB_TT_KeywordOrIdentifier) return false;
    uint32_t st;
    {
      const uint32_t first = p >> 8;
      if (first == KW_SELECT) st = FS_SEL0;
      else if (first == KW_INSERT) st = FS_INS0;
      else if (first == KW_CREATE) st = FS_CRE0;
      else if (WIDE && first == KW_WITH) st = FS_WITH0;
      else return false;
    }
    const uint32_t cap = nd.capacity();
    uint32_t t = 1, n = 0, sp = 0;
    uint32_t cur_start = 0, cur_kind = 0;  // the operand just completed (right-most subtree)
    uint32_t m0 = 0, m1 = 0;               // open interior nodes: outer (ROWS / COLDEF) and inner (clause / ROW / attribute)
    uint32_t ctx = C_SEL_ITEM, seen = 0, cnt = 0, width = 0, auxr = 0, jreg = 0;
    uint32_t qbase = 0;                    // first node of the query being parsed (wide: moves with subqueries)
#define FAST_STK(i_) stk[(int32_t)(i_) * stride]
#define FAST_EMIT(kind_, sub_, aux_, x_)          \
  do {                                            \
    if (n < cap) {                                \
      CNode c_;                                   \
      c_.kind = (uint8_t)(kind_);                 \
      c_.sub = (uint8_t)(sub_);                   \
      c_.aux = (uint16_t)(aux_);                  \
      c_.x = (x_);                                \
      nd.set(n, c_);                              \
    }                                             \
    n++;                                          \
  } while (0)
    for (;;) {
      // (a store beyond the range was skipped: the automaton redoes the statement.  Wide: the stack comes down from
      // the top of the same range; one turn of the loop emits / pushes at most three entries)
      if (WIDE ? (n + sp + 4u > cap) : (n > cap)) return false;
      p = tok.pair_at(t);  // the current token never lies beyond the statement's EOF token
      const uint32_t ty = p & 255u, kw = p >> 8;
      const uint32_t ti = fast_token_index(ty, kw);
      if (st == FS_X_OPER) {
        // ---------------- operators: token_power (mod.rs:1895-1927) + must_parse_expr_infix ----------------
        const uint32_t e = F->op[ti];
        const uint32_t spec = e >> 12;
        if (!WIDE && spec == SPEC_BAIL) return false;  // index access: the wide pass / the automaton
        const uint32_t power = e & 15u, op = (e >> 4) & 63u;
        // everything of equal or higher power on the stack is complete (left-associative)
        bool took_and = false;
        while (sp > 0) {
          const FastStackEntry top = FAST_STK(sp - 1);
          const uint32_t type = top.x & 15u;
          if (((top.x >> X_POWER) & 15u) < power) break;  // (brackets carry power 0: only a terminator gets past this)
          if (type == E_OP) {
            // BinaryOp{op, left, right}; refuse whatever simplify.rs would fold
            const uint32_t bop = (top.x >> X_OP) & 63u, lkind = (top.x >> X_LKIND) & 255u;
            if (bop == 9 || bop == 10) {  // simplified_eq / simplified_neq (simplify.rs): literal = literal folds
              if (is_literal(lkind) && is_literal(cur_kind)) {
                if (!WIDE || n < 2u || cur_start != n - 1u || top.y != n - 2u) return false;
                const uint32_t eq = lit_equal_simple(nd.get(n - 2u), nd.get(n - 1u));
                if (eq > 1u) return false;
                sp--;
                n -= 2u;
                cur_start = n;
                cur_kind = NUTDB_NK_LIT_BOOL;
                FAST_EMIT(NUTDB_NK_LIT_BOOL, ((bop == 9) == (eq == 1u)) ? 1 : 0, 0, NUTDB_CN_NOTOK);
                continue;
              }
            } else if (bop >= 11 && bop <= 13) {  // simplified_and / or / xor: a boolean literal operand folds
              if (lkind == NUTDB_NK_LIT_BOOL || cur_kind == NUTDB_NK_LIT_BOOL) {
                if (!WIDE || n > cap) return false;
                sp--;
                const uint32_t li = top.y;  // the left operand's first node; the right operand is [cur_start, n)
                if (lkind == NUTDB_NK_LIT_BOOL) {  // (the left operand is that one leaf: cur_start == li + 1)
                  const bool b = nd.get(li).sub != 0;
                  if (bop == 11 ? !b : (bop == 12 ? b : false)) {  // false AND x = false, true OR x = true
                    n = li + 1u;
                    cur_kind = NUTDB_NK_LIT_BOOL;
                  } else {  // the right operand alone; true XOR x = NOT x
                    remove_leaf(li, n);
                    if (bop == 13 && b) FAST_EMIT(NUTDB_NK_UNARY, 1, 0, li);
                    cur_kind = nd.get(n - 1u).kind;
                  }
                } else {  // the right operand is the boolean leaf n - 1
                  const bool b = nd.get(n - 1u).sub != 0;
                  if (bop == 11 ? !b : (bop == 12 ? b : false)) {  // x AND false = false, x OR true = true
                    n = li;
                    FAST_EMIT(NUTDB_NK_LIT_BOOL, b ? 1 : 0, 0, NUTDB_CN_NOTOK);
                    cur_kind = NUTDB_NK_LIT_BOOL;
                  } else {  // the left operand alone; x XOR true = NOT x
                    n -= 1u;
                    if (bop == 13 && b) FAST_EMIT(NUTDB_NK_UNARY, 1, 0, li);
                    cur_kind = nd.get(n - 1u).kind;
                  }
                }
                cur_start = li;
                continue;
              }
            }
            sp--;
            cur_start = top.y;
            cur_kind = NUTDB_NK_BINARY;
            FAST_EMIT(NUTDB_NK_BINARY, bop, 0, cur_start);
          } else if (type >= E_PAREN) {  // a bracket: the terminator belongs to it
            break;
          } else if (type == E_NOT) {  // simplified_not (simplify.rs): a boolean literal is flipped
            if (cur_kind == NUTDB_NK_LIT_BOOL) {
              if (!WIDE || n > cap || n == 0u) return false;
              CNode o = nd.get(n - 1u);
              o.sub = o.sub ? 0 : 1;
              nd.set(n - 1u, o);
              sp--;
              continue;
            }
            sp--;
            cur_kind = NUTDB_NK_UNARY;
            FAST_EMIT(NUTDB_NK_UNARY, 1, 0, cur_start);
          } else if (type == E_BITNOT) {  // UnaryOp{BitNot} (mod.rs:1290-1292): never folded
            sp--;
            cur_kind = NUTDB_NK_UNARY;
            FAST_EMIT(NUTDB_NK_UNARY, 0, 0, cur_start);
          } else if (type == E_BTW1) {  // `left BETWEEN x` must continue with AND (mod.rs:1445-1449)
            if (!(ty == NUTDB_TT_KeywordOrIdentifier && kw == KW_AND)) return false;
            FAST_STK(sp - 1).x = (top.x & ~15u) | E_BTW2;
            took_and = true;
            break;
          } else {  // E_BTW2: FnCall{Between | NotBetween, [left, x, y]}
            sp--;
            cur_start = top.y;
            cur_kind = NUTDB_NK_FNCALL;
            FAST_EMIT(NUTDB_NK_FNCALL, (top.x >> X_OP) & 63u, 0, cur_start);
          }
        }
        if (took_and) {
          t++;
          st = FS_X_OPND;  // the upper bound
          continue;
        }
        if (spec != SPEC_NONE) {
          if (WIDE && spec == SPEC_BAIL) {  // left[index] (mod.rs:1372-1382): BinaryOp{IndexAccess}, closed by `]`
            FAST_STK(sp) = FastStackEntry{E_INDEX, cur_start};
            sp++;
            t++;
            st = FS_X_OPND;
            continue;
          }
          const uint32_t p1 = tok.pair_at(t + 1);  // (the current token is a word, so t + 1 is at most the EOF token)
          const uint32_t kw1 = (p1 & 255u) == NUTDB_TT_KeywordOrIdentifier ? (p1 >> 8) : 0u;
          if (spec == SPEC_IS) {  // IS [NOT] NULL (mod.rs:1430-1438): simplified_is_null folds literals
            uint32_t sub = 2, used = 2;
            if (kw1 == KW_NOT) {
              const uint32_t p2 = tok.pair_at(t + 2);
              if (!((p2 & 255u) == NUTDB_TT_KeywordOrIdentifier && (p2 >> 8) == KW_NULL)) return false;
              sub = 3;
              used = 3;
            } else if (kw1 != KW_NULL) {
              return false;
            }
            if (is_literal(cur_kind)) {  // simplified_is_null / _is_not_null: a literal operand folds
              if (!WIDE || cur_start != n - 1u) return false;
              const bool isnull = cur_kind == NUTDB_NK_LIT_NULL;
              n -= 1u;
              cur_kind = NUTDB_NK_LIT_BOOL;
              FAST_EMIT(NUTDB_NK_LIT_BOOL, ((sub == 2) == isnull) ? 1 : 0, 0, NUTDB_CN_NOTOK);
              t += used;
              continue;
            }
            cur_kind = NUTDB_NK_UNARY;
            FAST_EMIT(NUTDB_NK_UNARY, sub, 0, cur_start);
            t += used;
            continue;  // still after an operand
          }
          if (sp >= DEPTH) return false;
          if (spec == SPEC_BETWEEN) {
            FAST_STK(sp) = FastStackEntry{E_BTW1 | (P_Between << X_POWER) | (3u << X_OP), cur_start};
            t += 1;
          } else {  // NOT IN / LIKE / ILIKE / BETWEEN (mod.rs:1399-1427); NOT EXISTS goes to the automaton
            if (kw1 == KW_BETWEEN) FAST_STK(sp) = FastStackEntry{E_BTW1 | (P_Between << X_POWER) | (4u << X_OP), cur_start};
            else if (kw1 == KW_IN || kw1 == KW_LIKE || kw1 == KW_ILIKE)
              FAST_STK(sp) = FastStackEntry{E_OP | (P_Comparison << X_POWER) | ((kw1 == KW_IN ? 19u : kw1 == KW_LIKE ? 15u : 17u) << X_OP) |
                                                (cur_kind << X_LKIND), cur_start};
            else return false;
            t += 2;
          }
          sp++;
          st = FS_X_OPND;
          continue;
        }
        if (power != P_Terminator) {
          if (sp >= DEPTH) return false;
          FAST_STK(sp) = FastStackEntry{E_OP | (power << X_POWER) | (op << X_OP) | (cur_kind << X_LKIND), cur_start};
          sp++;
          t++;
          st = FS_X_OPND;  // right operand
          continue;
        }
        // a terminator.  Inside a bracket opened by this expression it belongs to the bracket; a subquery's frame or a
        // pending set operation is not such a bracket: the expression of the inner query is complete and the grammar
        // table decides
        if (sp != 0 && !(WIDE && (FAST_STK(sp - 1).x & 15u) >= E_SUBQ)) {
          const FastStackEntry br0 = FAST_STK(sp - 1);
          const uint32_t btype = br0.x & 15u;
          const uint32_t k = ty == NUTDB_TT_KeywordOrIdentifier ? kw : 0u;
          if (btype == E_CASE) {  // must_parse_case_when_body (mod.rs:1585-1618)
            const uint32_t at = (br0.x >> 8) & 3u;
            uint32_t next;
            if ((at == 0 || at == 2) && k == KW_WHEN) next = 1;
            else if (at == 1 && k == KW_THEN) next = 2;
            else if (at == 2 && k == KW_ELSE) next = 3;
            else if (at >= 2 && k == KW_END) {  // without ELSE the default is NULL
              if (at == 2) FAST_EMIT(NUTDB_NK_LIT_NULL, 0, 0, NUTDB_CN_NOTOK);
              sp--;
              t++;
              cur_start = br0.y;
              cur_kind = NUTDB_NK_FNCALL;
              FAST_EMIT(NUTDB_NK_FNCALL, (br0.x >> 10) & 7u, 0, br0.y);
              continue;
            } else {
              return false;
            }
            FAST_STK(sp - 1).x = (br0.x & ~(3u << 8)) | (next << 8);
            t++;
            st = FS_X_OPND;
            continue;
          }
          if (WIDE) {
            if (btype == E_IF) {  // must_parse_if_body (mod.rs:1571-1582): IF c THEN a ELSE b END = FnName::If
              const uint32_t at = (br0.x >> 8) & 3u;
              if ((at == 0 && k == KW_THEN) || (at == 1 && k == KW_ELSE)) {
                FAST_STK(sp - 1).x = (br0.x & ~(3u << 8)) | ((at + 1u) << 8);
                t++;
                st = FS_X_OPND;
                continue;
              }
              if (at == 2 && k == KW_END) {
                sp--;
                t++;
                cur_start = br0.y;
                cur_kind = NUTDB_NK_FNCALL;
                FAST_EMIT(NUTDB_NK_FNCALL, 0, 0, br0.y);
                continue;
              }
              return false;
            }
            if (btype == E_MAP) {  // must_parse_map (mod.rs:1558-1568): key : value {, key : value} }
              const uint32_t at = (br0.x >> 8) & 1u;
              if ((at == 0 && ty == NUTDB_TT_Colon) || (at == 1 && ty == NUTDB_TT_Comma)) {
                FAST_STK(sp - 1).x = br0.x ^ (1u << 8);
                t++;
                st = FS_X_OPND;
                continue;
              }
              if (at == 1 && ty == NUTDB_TT_RBrace) {
                sp--;
                t++;
                cur_start = br0.y;
                cur_kind = NUTDB_NK_COLLECTION;
                FAST_EMIT(NUTDB_NK_COLLECTION, 1, 0, br0.y);
                continue;
              }
              return false;
            }
            if (btype == E_BRACKET) {  // [items] (mod.rs:1247-1251): Collection{Array}
              if (ty == NUTDB_TT_Comma) {
                t++;
                st = FS_X_OPND;
                continue;
              }
              if (ty != NUTDB_TT_RBracket) return false;
              sp--;
              t++;
              cur_start = br0.y;
              cur_kind = NUTDB_NK_COLLECTION;
              FAST_EMIT(NUTDB_NK_COLLECTION, 2, 0, br0.y);
              continue;
            }
            if (btype == E_INDEX) {  // left[index]: exactly one expression, then `]`
              if (ty != NUTDB_TT_RBracket) return false;
              sp--;
              t++;
              cur_start = br0.y;
              cur_kind = NUTDB_NK_BINARY;
              FAST_EMIT(NUTDB_NK_BINARY, 20, 0, br0.y);
              continue;
            }
          }
          if (ty == NUTDB_TT_Comma) {
            if (((br0.x >> X_COUNT) & X_COUNT_MAX) == X_COUNT_MAX) return false;  // (the item counter is full)
            FAST_STK(sp - 1).x += 1u << X_COUNT;
            t++;
            st = FS_X_OPND;  // next item
            continue;
          }
          if (ty != NUTDB_TT_RParen) return false;
          const FastStackEntry br = FAST_STK(--sp);
          t++;
          if ((br.x & 15u) == E_CALL) {
            FAST_EMIT(NUTDB_NK_FNCALL, 7, 0, br.y);
            cur_start = br.y;
            cur_kind = NUTDB_NK_FNCALL;
          } else if ((br.x >> X_COUNT) != 0u) {  // one item in parentheses is the item itself (mod.rs:1236-1242)
            FAST_EMIT(NUTDB_NK_COLLECTION, 0, 0, br.y);
            cur_start = br.y;
            cur_kind = NUTDB_NK_COLLECTION;
          }
          continue;
        }
        st = FS_AFTER + ctx;  // the expression is complete: the same token decides what follows it
      }
      const uint32_t cls = F->cls[ti];
      if (WIDE && !NUTDB_WIDE_LITE && st == FS_END_SEL && cls != FC_SETOP) {
        // the query expression ends here: the set operations still waiting for their right query get it
        // (must_parse_query_tdop, mod.rs:243-276), innermost first
        while (sp > 0 && (FAST_STK(sp - 1).x & 15u) == E_UNION) {
          const FastStackEntry u = FAST_STK(--sp);
          FAST_EMIT(NUTDB_NK_QUERY_UNION, (u.x >> 8) & 3u, 0, u.y);
          qbase = u.y;  // (the whole expression now starts where its left-most query does)
        }
        // CREATE VIEW .. AS query: the statement ends with the query (no parenthesis): ViewDefinition + CreateStmt
        if (sp >= 2u && (FAST_STK(sp - 1).x & (15u | SUBQ_VIEW)) == (E_SUBQ | SUBQ_VIEW) && (cls == FC_EOF || cls == FC_SEMI)) {
          sp -= 2u;
          qbase = 0u;
          FAST_EMIT(NUTDB_NK_VIEWDEF, 0, 0, 0u);
          FAST_EMIT(NUTDB_NK_STMT_CREATE, 0, auxr, 0u);
          st = FS_FINAL;
        }
      }
      const uint32_t ri = F->trans[st][cls];
      const uint32_t lo = F->rec_lo[ri];
      const uint32_t act = lo & 15u;
      if (act == FA_STEP) {
        const uint32_t hi = F->rec_hi[ri];
        const uint32_t kind = (lo >> 15) & 255u;
        if (hi & FAST_HI_UNCOMMON) {
          const uint32_t look = (hi >> 8) & 3u;
          if (look) {  // (the current token is not EOF here, so t + 1 exists)
            const uint32_t ty1 = tok.pair_at(t + 1) & 255u;
            if ((look != FL_NOLP && ty1 == NUTDB_TT_Dot) || (look >= FL_NODOT_NOLP && ty1 == NUTDB_TT_LParen)) return false;
          }
          const uint32_t check = (hi >> 5) & 7u;
          if (check == FK_STR) {
            // (kw byte 1: the lexer saw no backslash-u escape in the literal, nothing can be rejected)
            if (kw != 1u && !string_ok(t, ty == NUTDB_TT_EscapedSQStringLiteral ? '\'' : '"')) return false;
          } else if (check == FK_FNCALL) {  // the indexer of INDEX name f(..) must be a function call (mod.rs:923-931)
            if (cur_kind != NUTDB_NK_FNCALL) return false;
          } else if (check != FK_NONE) {
            if (!int_ok(ty, kw, check - FK_INT_W0)) return false;
          }
          const uint32_t bit = (hi >> 15) & 255u;
          if (seen & bit) return false;  // Conflicts: the automaton reports it
          seen |= bit;
          if (hi & (1u << 23)) seen &= ~3u;
          if (hi & (1u << 24)) cnt++;
          if (hi & (1u << 25)) auxr = 1;
          if (hi & (1u << 10)) ctx = (hi >> 11) & 15u;
          if (hi & (1u << 26)) jreg = (hi >> 27) & 7u;
          if (hi & (1u << 30)) n--;
        }
        if (hi & 1u) m0 = n;
        if (hi & 2u) m1 = n;
        if (hi & 16u) {
          cur_start = n;
          cur_kind = kind;
        }
        const uint32_t em = (lo >> 12) & 7u;
        if (em != FE_NONE) {
          const uint32_t x = em == FE_LEAF_TOK ? t : em == FE_LEAF_NOTOK ? NUTDB_CN_NOTOK : em == FE_NODE_M0 ? m0
                             : em == FE_NODE_M1 ? m1 : (WIDE ? qbase : 0u);
          if (n < cap) nd.set_raw(n, F->rec_hdr[ri] | (((lo >> 28) & auxr) << 16) | ((lo & (1u << 29)) ? jreg << 8 : 0u), x);
          n++;
        }
        t += (lo >> 11) & 1u;
        if (hi & 4u) m0 = n;
        if (hi & 8u) m1 = n;
        st = (lo >> 4) & 127u;
      } else if (act == FA_IDENT) {
        // an identifier-like word in operand position: plain, qualified (mod.rs:1506-1523) or a call (:1303-1308, :1538-1556)
        const uint32_t p1 = tok.pair_at(t + 1);  // (the current token is a word, so t + 1 is at most the EOF token)
        const uint32_t ty1 = p1 & 255u;
        if (ty1 == NUTDB_TT_LParen && ty == NUTDB_TT_KeywordOrIdentifier) {
          const uint32_t p2 = tok.pair_at(t + 2);
          if ((p2 & 255u) == NUTDB_TT_KeywordOrIdentifier && ((p2 >> 8) == KW_SELECT || (p2 >> 8) == KW_WITH)) {
            // name(select ..): the subquery is the only argument, its `)` the call's (try_parse_fn_call_args, mod.rs:1538-1556)
            if constexpr (WIDE) {
              if ((p2 >> 8) == KW_WITH) return false;
              FAST_EMIT(NUTDB_NK_FN_NAME, 0, 0, t);
              FAST_STK(sp) = FastStackEntry{m0, m1};
              FAST_STK(sp + 1) = FastStackEntry{E_SUBQ | (ctx << 8) | (jreg << 12) | SUBQ_CALL, qbase};
              sp += 2;
              qbase = n;
              m0 = n;
              m1 = n;
              ctx = C_SEL_ITEM;
              t += 3;  // name ( SELECT
              st = FS_SEL0;
              continue;
            } else {
              return false;
            }
          }
          const uint32_t m = n;
          FAST_EMIT(NUTDB_NK_FN_NAME, 0, 0, t);
          if ((p2 & 255u) == NUTDB_TT_RParen) {
            FAST_EMIT(NUTDB_NK_FNCALL, 7, 0, m);
            cur_start = m;
            cur_kind = NUTDB_NK_FNCALL;
            t += 3;
            st = FS_X_OPER;
          } else {
            if (sp >= DEPTH) return false;
            FAST_STK(sp) = FastStackEntry{E_CALL, m};
            sp++;
            t += 2;
            st = FS_X_OPND;  // first argument
          }
        } else if (ty1 == NUTDB_TT_Dot) {
          const uint32_t ty2 = tok.pair_at(t + 2) & 255u;
          if (!(ident_string(ty2) || ty2 == NUTDB_TT_Mul)) return false;
          cur_start = n;
          cur_kind = NUTDB_NK_IDENT;
          FAST_EMIT(NUTDB_NK_QUAL, 0, 0, t);
          FAST_EMIT(NUTDB_NK_IDENT, ty2 == NUTDB_TT_Mul ? 1 : 0, 1, t + 2);
          t += 3;
          st = FS_X_OPER;
        } else {
          cur_start = n;
          cur_kind = NUTDB_NK_IDENT;
          FAST_EMIT(NUTDB_NK_IDENT, 0, 0, t);
          t++;
          st = FS_X_OPER;
        }
      } else if (act == FA_OPEN) {  // (mod.rs:1229-1246)
        const uint32_t p1 = tok.pair_at(t + 1);
        if ((p1 & 255u) == NUTDB_TT_KeywordOrIdentifier && ((p1 >> 8) == KW_SELECT || (p1 >> 8) == KW_WITH)) {
          // a subquery (must_parse_subquery, mod.rs:206-241): narrow -> the wide pass; WITH -> the automaton
          if constexpr (WIDE) {
            if ((p1 >> 8) == KW_WITH) return false;
            FAST_STK(sp) = FastStackEntry{m0, m1};
            FAST_STK(sp + 1) = FastStackEntry{E_SUBQ | (ctx << 8) | (jreg << 12), qbase};
            sp += 2;
            qbase = n;
            m0 = n;
            m1 = n;
            ctx = C_SEL_ITEM;
            t += 2;  // `(` and SELECT
            st = FS_SEL0;
            continue;
          } else {
            return false;
          }
        }
        if (sp >= DEPTH) return false;
        FAST_STK(sp) = FastStackEntry{E_PAREN, n};
        sp++;
        t++;
        st = FS_X_OPND;
      } else if (act == FA_CASE) {  // CASE WHEN .. = FnName::MultiIf (1), CASE scrutinee WHEN .. = FnName::CaseWhen (2)
        if (sp >= DEPTH) return false;
        const uint32_t p1 = tok.pair_at(t + 1);
        const bool multi = (p1 & 255u) == NUTDB_TT_KeywordOrIdentifier && (p1 >> 8) == KW_WHEN;
        FAST_STK(sp) = FastStackEntry{E_CASE | ((multi ? 1u : 0u) << 8) | ((multi ? 1u : 2u) << 10), n};
        sp++;
        t += multi ? 2u : 1u;
        st = FS_X_OPND;
      } else if (act == FA_PUSH) {
        // pushes the stack entry named in the record's kind field and goes on with an operand.  Prefix NOT (both
        // passes) applies to the next PREFIX expression only (mod.rs:1294-1296): `not a = b` is `(not a) = b` -- so its
        // entry (power 15) is completed by whatever token comes next; prefix ~ likewise.  Wide pass only: IF c THEN a
        // ELSE b END (mod.rs:1297-1299), [items] (at least one: must_parse_expr_list), {key : value, ..}.
        const uint32_t entry = (lo >> 15) & 255u;
        if (sp >= DEPTH || (!WIDE && (entry & 15u) != E_NOT)) return false;
        FAST_STK(sp) = FastStackEntry{entry, n};
        sp++;
        t++;
        st = FS_X_OPND;
      } else if (act == FA_NEG) {  // only a literal may follow a prefix minus here (mod.rs:1259-1269)
        const uint32_t p1 = tok.pair_at(t + 1);
        const uint32_t ty1 = p1 & 255u;
        cur_start = n;
        if (ty1 == NUTDB_TT_FloatLiteral) {
          cur_kind = NUTDB_NK_LIT_FLOAT;
          FAST_EMIT(NUTDB_NK_LIT_FLOAT, 1, 0, t + 1);
        } else if (ty1 == NUTDB_TT_IntegerLiteral || ty1 == NUTDB_TT_HexLiteral) {
          if (!int_ok(ty1, p1 >> 8, 2)) return false;
          cur_kind = NUTDB_NK_LIT_INT;
          FAST_EMIT(NUTDB_NK_LIT_INT, 1, ty1 == NUTDB_TT_HexLiteral ? 1 : 0, t + 1);
        } else {
          return false;
        }
        t += 2;
        st = FS_X_OPER;
      } else if (act == FA_DTYPE) {  // must_parse_datatype (mod.rs:1688-1797) without Enum / Tuple / Map
        const uint32_t i = kw - KW_INT8;
        const uint32_t ty1 = tok.pair_at(t + 1) & 255u;
        if (WIDE && !NUTDB_WIDE_LITE && i == 27) {  // Enum('a' [= n], ..) (must_parse_enum_binds, mod.rs:1799-1813)
          if (ty1 != NUTDB_TT_LParen || !parse_enum(t, n, sp, cap)) return false;
          st = sp ? (uint32_t)FS_DT_END : (uint32_t)FS_COL_ATTRS;
        } else if (i == 26 || i == 30 || i == 31) {  // Array / Dictionary / Nullable (inner)
          if (ty1 != NUTDB_TT_LParen || sp >= DEPTH) return false;
          FAST_STK(sp) = FastStackEntry{E_DT | ((i == 26 ? 0u : (i == 30 ? 4u : 5u)) << 4), n};
          sp++;
          t += 2;  // the inner type follows (state FS_DT stays)
        } else {
          if (i > 25) return false;
          if (i == 16 || i == 17 || i == 21 || (i == 22 && ty1 == NUTDB_TT_LParen)) {  // Type(n)
            if (ty1 != NUTDB_TT_LParen) return false;
            const uint32_t p2 = tok.pair_at(t + 2);
            const uint32_t ty2 = p2 & 255u;
            if (ty2 != NUTDB_TT_IntegerLiteral && ty2 != NUTDB_TT_HexLiteral) return false;
            if (!int_ok(ty2, p2 >> 8, (i == 16 || i == 17) ? 0u : 1u)) return false;
            if ((tok.pair_at(t + 3) & 255u) != NUTDB_TT_RParen) return false;
            const uint32_t m = n;
            FAST_EMIT(NUTDB_NK_NUM, 0, ty2 == NUTDB_TT_HexLiteral ? 1 : 0, t + 2);
            FAST_EMIT(NUTDB_NK_DT_PARAM, i, 0, m);
            t += 4;
          } else {
            FAST_EMIT(NUTDB_NK_DT_SCALAR, i, 0, NUTDB_CN_NOTOK);
            t++;
          }
          st = sp ? (uint32_t)FS_DT_END : (uint32_t)FS_COL_ATTRS;
        }
      } else if (act == FA_DTEND) {  // the closing parentheses of compound types
        if (ty != NUTDB_TT_RParen || sp == 0) return false;
        const FastStackEntry d = FAST_STK(--sp);
        if ((d.x & 15u) != E_DT) return false;
        FAST_EMIT(NUTDB_NK_DT_COMPOUND, d.x >> 4, 0, d.y);
        t++;
        st = sp ? (uint32_t)FS_DT_END : (uint32_t)FS_COL_ATTRS;
      } else if (act == FA_ROWEND) {  // `)` of a VALUES row: must_parse_insert_rows (mod.rs:636-670)
        cnt++;
        if (width == 0) width = cnt;
        else if (cnt != width) return false;  // Conflicts: the automaton reports it
        cnt = 0;
        FAST_EMIT(NUTDB_NK_ROW, 0, 0, m1);
        t++;
        st = FS_INS_AFTER_ROW;
      } else if (act == FA_ACCEPT) {
        if (sp != 0 || n > cap) return false;
        res.status = NUTDB_ST_OK;
        res.node_count = n;
        res.tok_used = t + 1;
        res.err_code = 0;
        res.err_has_pos = false;
        res.err_pos = res.err_a = res.err_b = res.err_c = 0;
        return true;
      } else if (WIDE && !NUTDB_WIDE_LITE && act == FA_SETOP) {
        // UNION ALL | UNION DISTINCT | INTERSECT | EXCEPT behind a query body (mod.rs:250-267): precedence climbing over
        // UnionTypePower (Except < Union < Intersect), left-associative like the expression operators
        uint32_t power, utype, used = 1;
        if (kw == KW_UNION) {
          const uint32_t p1 = tok.pair_at(t + 1);
          const uint32_t k1 = (p1 & 255u) == NUTDB_TT_KeywordOrIdentifier ? (p1 >> 8) : 0u;
          if (k1 == KW_ALL) utype = 0;
          else if (k1 == KW_DISTINCT) utype = 1;
          else return false;
          power = U_Union;
          used = 2;
        } else if (kw == KW_INTERSECT) {
          power = U_Intersect;
          utype = 2;
        } else {
          power = U_Except;
          utype = 3;
        }
        uint32_t left = qbase;
        while (sp > 0 && (FAST_STK(sp - 1).x & 15u) == E_UNION && ((FAST_STK(sp - 1).x >> X_POWER) & 15u) >= power) {
          const FastStackEntry u = FAST_STK(--sp);
          FAST_EMIT(NUTDB_NK_QUERY_UNION, (u.x >> 8) & 3u, 0, u.y);
          left = u.y;
        }
        const uint32_t pn = tok.pair_at(t + used);  // the right query: a plain SELECT here (`(` / WITH: the automaton)
        if (!((pn & 255u) == NUTDB_TT_KeywordOrIdentifier && (pn >> 8) == KW_SELECT)) return false;
        FAST_STK(sp) = FastStackEntry{E_UNION | (power << X_POWER) | (utype << 8), left};
        sp++;
        qbase = n;
        m0 = n;
        m1 = n;
        ctx = C_SEL_ITEM;
        t += used + 1u;
        st = FS_SEL0;
      } else if (WIDE && act == FA_INTERVAL) {  // INTERVAL n unit (must_parse_interval, mod.rs:1489-1503)
        const uint32_t p1 = tok.pair_at(t + 1);
        const uint32_t ty1 = p1 & 255u;
        if (ty1 != NUTDB_TT_IntegerLiteral && ty1 != NUTDB_TT_HexLiteral) return false;
        if (!int_ok(ty1, p1 >> 8, 1)) return false;
        const uint32_t p2 = tok.pair_at(t + 2);
        const uint32_t k2 = (p2 & 255u) == NUTDB_TT_KeywordOrIdentifier ? (p2 >> 8) : 0u;
        uint32_t unit;
        if (k2 == KW_SECOND) unit = 0;
        else if (k2 == KW_MINUTE) unit = 1;
        else if (k2 == KW_HOUR) unit = 2;
        else if (k2 == KW_DAY) unit = 3;
        else if (k2 == KW_MONTH) unit = 4;
        else if (k2 == KW_YEAR) unit = 5;
        else return false;
        cur_start = n;
        cur_kind = NUTDB_NK_LIT_INTERVAL;
        FAST_EMIT(NUTDB_NK_LIT_INTERVAL, unit, ty1 == NUTDB_TT_HexLiteral ? 1 : 0, t + 1);
        t += 3;
        st = FS_X_OPER;
      } else if (WIDE && act == FA_SRC_SUBQ) {  // FROM (select ..) [AS alias] (must_parse_query_source, mod.rs:546-569);
                                                // WITH name AS (select ..) (must_parse_query_clause_with, mod.rs:327-347)
        const uint32_t p1 = tok.pair_at(t + 1);
        if (!((p1 & 255u) == NUTDB_TT_KeywordOrIdentifier && (p1 >> 8) == KW_SELECT)) return false;
        FAST_STK(sp) = FastStackEntry{m0, m1};
        const uint32_t what = (lo >> 15) & 255u;  // 0: a query source, 1: a common table expression, 2: the query of CREATE VIEW .. AS
        if (what == 2u && !(seen & 1u)) return false;  // (AS needs the UPDATE BY first, mod.rs:824-828)
        FAST_STK(sp + 1) = FastStackEntry{E_SUBQ | (ctx << 8) | (jreg << 12) | (what == 0u ? SUBQ_SOURCE : what == 1u ? SUBQ_CTE : SUBQ_VIEW), qbase};
        sp += 2;
        qbase = n;
        m0 = n;
        m1 = n;
        ctx = C_SEL_ITEM;
        t += 2;
        st = FS_SEL0;
      } else if (WIDE && act == FA_SUBQ_END) {  // `)` right after a query body: the subquery's own parenthesis
        if (sp < 2u || (FAST_STK(sp - 1).x & (15u | SUBQ_VIEW)) != E_SUBQ) return false;  // (a view's query has no parenthesis)
        const FastStackEntry f1 = FAST_STK(sp - 1), f0 = FAST_STK(sp - 2);
        sp -= 2;
        cur_start = qbase;
        cur_kind = NUTDB_NK_QUERY_BODY;
        ctx = (f1.x >> 8) & 15u;
        jreg = (f1.x >> 12) & 7u;
        qbase = f1.y;
        m0 = f0.x;
        m1 = f0.y;
        t++;
        st = FS_X_OPER;  // the outer expression goes on behind it
        if (f1.x & SUBQ_CALL) {  // name(select ..): FnCall{Others, [subquery]} over the name in front of it
          cur_start -= 1u;
          cur_kind = NUTDB_NK_FNCALL;
          FAST_EMIT(NUTDB_NK_FNCALL, 7, 0, cur_start);
        } else if (f1.x & SUBQ_SOURCE) {
          st = FS_SRC2B;  // a query source: [AS alias], then the clauses; an operator behind it goes to the automaton
        } else if (f1.x & SUBQ_CTE) {
          st = FS_WITH_SEP;  // a common table expression: `,` and the next one, or SELECT
        }
      } else {
        return false;
      }
    }
#undef FAST_EMIT
#undef FAST_STK
  }
};

}  // namespace npar
