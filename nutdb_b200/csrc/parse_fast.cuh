// parse_fast.cuh -- straight-line parser for the statement shapes that dominate query logs.
//
// The bytecode automaton in parse_core.cuh is exact for the whole grammar but pays an
// interpreter's price (fetch/dispatch/stack per grammar step) and its lanes diverge on every
// dispatch.  This file parses the COMMON shapes directly:
//
//   SELECT items [FROM name [AS a]] [WHERE e] [GROUP BY items] [HAVING e] [ORDER BY item [DESC],..]
//          [LIMIT n [, m | OFFSET m] [WITH TIES]]
//   INSERT INTO name [(names)] VALUES (exprs) {, (exprs)}
//   CREATE TABLE [IF NOT EXISTS] name (name type [DEFAULT e | COMMENT s].. ,..)
//          [PRIMARY KEY es | ORDER BY es | PARTITION BY e | COMMENT s]..
//
// with expressions over identifiers, literals, the binary operators, AND/OR/XOR, IN/LIKE/ILIKE,
// parenthesised groups / tuples and function calls (operator-precedence parsing with an explicit
// operator stack -- the iterative form of must_parse_expr_tdop, reference mod.rs:1209-1220: an
// operator is reduced when one of equal or lower power arrives, so every operator is
// left-associative exactly as in the reference).
//
// It is ALL-OR-NOTHING: on anything outside that subset -- any error, any construct that needs
// constant folding (simplify.rs), literal validation beyond a length check, joins, set
// operations, subqueries, CASE/IF/NOT/IS/BETWEEN, arrays, maps ... -- try_parse() returns false
// without side effects the caller keeps, and the statement is parsed from scratch by the exact
// automaton.  For a statement it accepts, it emits precisely the nodes the automaton would.
#pragma once
#include "parse_core.cuh"

namespace npar {

// Lookup tables of the expression loop: what an operand token is, and the infix power / operator of a token
// (token_power, mod.rs:1895-1927).  Lanes holding different token types then execute the SAME instructions.
struct FastTables {
  uint32_t opnd[48];   // by token type: class | kind << 8 | sub << 16 | aux << 24
  uint16_t optok[48];  // by token type: power | op << 4 | bail << 12
  uint16_t opkw[128];  // by keyword id: same encoding
};
enum : uint32_t { FO_BAIL = 0, FO_LEAF = 1, FO_WORD = 2, FO_DELIM = 3, FO_MINUS = 4, FO_PLUS = 5, FO_LPAREN = 6, FO_INT = 7,
                  FO_ESTR = 8 };
NUTDB_HD uint32_t fast_opnd_entry(uint32_t ty) {
  switch (ty) {
    case NUTDB_TT_KeywordOrIdentifier: return FO_WORD;
    case NUTDB_TT_DelimitedIdentifier: return FO_DELIM;
    case NUTDB_TT_Mul: return FO_LEAF | (NUTDB_NK_IDENT << 8) | (1u << 16);
    case NUTDB_TT_RawStringLiteral: return FO_LEAF | (NUTDB_NK_LIT_STR << 8);
    case NUTDB_TT_EscapedSQStringLiteral: return FO_ESTR | (NUTDB_NK_LIT_STR << 8) | (1u << 16);
    case NUTDB_TT_EscapedDQStringLiteral: return FO_ESTR | (NUTDB_NK_LIT_STR << 8) | (2u << 16);
    case NUTDB_TT_FloatLiteral: return FO_LEAF | (NUTDB_NK_LIT_FLOAT << 8);
    case NUTDB_TT_IntegerLiteral: return FO_INT | (NUTDB_NK_LIT_INT << 8);
    case NUTDB_TT_HexLiteral: return FO_INT | (NUTDB_NK_LIT_INT << 8) | (1u << 24);
    case NUTDB_TT_Minus: return FO_MINUS;
    case NUTDB_TT_Plus: return FO_PLUS;
    case NUTDB_TT_LParen: return FO_LPAREN;
    default: return FO_BAIL;
  }
}
NUTDB_HD uint16_t fast_optok_entry(uint32_t ty) {
  uint32_t power = P_Terminator, op = 0, bail = 0;
  switch (ty) {
    case NUTDB_TT_Eq: power = P_Comparison; op = 9; break;
    case NUTDB_TT_NotEq: power = P_Comparison; op = 10; break;
    case NUTDB_TT_Gt: power = P_Comparison; op = 5; break;
    case NUTDB_TT_Lt: power = P_Comparison; op = 6; break;
    case NUTDB_TT_GtEq: power = P_Comparison; op = 7; break;
    case NUTDB_TT_LtEq: power = P_Comparison; op = 8; break;
    case NUTDB_TT_BitOr: power = P_BitOr; op = 21; break;
    case NUTDB_TT_BitXor: power = P_BitXor; op = 23; break;
    case NUTDB_TT_BitAnd: power = P_BitAnd; op = 22; break;
    case NUTDB_TT_BitLShift: power = P_BitShift; op = 24; break;
    case NUTDB_TT_BitRShift: power = P_BitShift; op = 25; break;
    case NUTDB_TT_Plus: power = P_PlusMinus; op = 0; break;
    case NUTDB_TT_Minus: power = P_PlusMinus; op = 1; break;
    case NUTDB_TT_Mul: power = P_MulDivMod; op = 2; break;
    case NUTDB_TT_Div: power = P_MulDivMod; op = 3; break;
    case NUTDB_TT_Mod: power = P_MulDivMod; op = 4; break;
    case NUTDB_TT_LBracket: bail = 1; break;  // index access
    default: break;
  }
  return (uint16_t)(power | (op << 4) | (bail << 12));
}
NUTDB_HD uint16_t fast_opkw_entry(uint32_t kw) {
  uint32_t power = P_Terminator, op = 0, bail = 0;
  switch (kw) {
    case KW_AND: power = P_And; op = 11; break;
    case KW_OR: power = P_Or; op = 12; break;
    case KW_XOR: power = P_Xor; op = 13; break;
    case KW_IN: power = P_Comparison; op = 18; break;
    case KW_LIKE: power = P_Comparison; op = 14; break;
    case KW_ILIKE: power = P_Comparison; op = 16; break;
    case KW_NOT: case KW_IS: case KW_BETWEEN: bail = 1; break;
    default: break;
  }
  return (uint16_t)(power | (op << 4) | (bail << 12));
}
NUTDB_HD void fast_tables_fill(FastTables& F, uint32_t i) {  // entry i of every table (i < 128)
  if (i < 48) {
    F.opnd[i] = fast_opnd_entry(i);
    F.optok[i] = fast_optok_entry(i);
  }
  F.opkw[i] = fast_opkw_entry(i);
}

template <class Tok, class Nodes, class Text>
struct FastParser {
  const FastTables& F;
  Tok& tok;
  Nodes& nd;
  Text& text;
  uint32_t t = 0, n = 0, cap;
  uint8_t ty = 0, kw = 0;  // the current token (index t), held in registers
  // the operand being built (right-most subtree)
  uint32_t cur_start = 0;
  uint8_t cur_kind = 0;
  // operator / bracket stack
  enum : uint32_t { E_OP = 0, E_PAREN = 1, E_CALL = 2 };
  static const uint32_t DEPTH = 12;
  uint32_t einfo[DEPTH];   // type | power << 2 | op << 6 | left kind << 12 | item count << 20
  uint32_t emark[DEPTH];   // E_OP: start of the left operand; brackets: node count at the opening
  uint32_t sp = 0;
  // where the expression being parsed sits in its statement (one shared expression loop: lanes of a
  // warp that are in different clauses still execute the same code)
  enum : uint8_t { C_SEL_ITEM, C_WHERE, C_GROUP_ITEM, C_HAVING, C_ORDER_ITEM, C_INS_VALUE, C_COL_DEFAULT, C_TBL_PK_ITEM,
                   C_TBL_ORDER_ITEM, C_TBL_PART };
  enum : uint32_t { R_BAIL = 0, R_EXPR = 1, R_DONE = 2 };
  uint8_t ctx = 0;
  uint32_t m0 = 0, m1 = 0;          // open interior nodes: outer (ROWS / COLDEF) and inner (clause / ROW / attribute)
  uint32_t width = 0, w = 0, row = 0, seen = 0, tseen = 0, aux = 0;

  NUTDB_HD FastParser(const FastTables& ft, Tok& tk, Nodes& nodes, Text& tx)
      : F(ft), tok(tk), nd(nodes), text(tx), cap(nodes.capacity()) {}

  NUTDB_HD void load() {
    ty = tok.type_at(t);
    kw = tok.kw_at(t);
  }
  NUTDB_HD void adv() {
    t++;
    load();
  }
  NUTDB_HD void adv(uint32_t k) {
    t += k;
    load();
  }
  NUTDB_HD bool emit(uint8_t kind, uint8_t sub, uint16_t ax, uint32_t x) {
    if (n >= cap) return false;
    CNode c;
    c.kind = kind;
    c.sub = sub;
    c.aux = ax;
    c.x = x;
    nd.set(n++, c);
    return true;
  }
  NUTDB_HD static bool is_literal(uint8_t kind) { return kind >= NUTDB_NK_LIT_INT && kind <= NUTDB_NK_LIT_INTERVAL; }
  NUTDB_HD bool is_kw(uint32_t k) const { return ty == NUTDB_TT_KeywordOrIdentifier && kw == k; }
  NUTDB_HD bool next_is_kw(uint32_t d, uint32_t k) { return tok.type(t + d) == NUTDB_TT_KeywordOrIdentifier && tok.kw(t + d) == k; }
  NUTDB_HD static bool ident_string(uint8_t y) {  // must_parse_identifier_string (mod.rs:1682)
    return y == NUTDB_TT_KeywordOrIdentifier || y == NUTDB_TT_DelimitedIdentifier;
  }
  // integer_from_str! cannot fail for 1..safe digits; the lexer stores min(len, 255) in the kw byte
  NUTDB_HD static bool int_ok(uint8_t y, uint32_t len, uint32_t width_) {
    const bool hex = y == NUTDB_TT_HexLiteral;
    const uint32_t safe = width_ == 0 ? 2u : width_ == 1 ? (hex ? 16u : 19u) : (hex ? 32u : 38u);
    return len >= 1 && len <= safe;
  }
  // an escaped string literal can only be rejected through a backslash-u escape (literal.rs:70-88)
  NUTDB_HD bool string_ok(uint32_t i, uint8_t y) {
    if (y == NUTDB_TT_RawStringLiteral) return true;
    const uint32_t s = tok.start(i), e = tok.end(i);
    for (uint32_t p = s; p + 1 < e; p++)
      if (text.byte(p) == '\\' && text.byte(p + 1) == 'u') return false;
    return true;
  }
  NUTDB_HD static uint8_t str_sub(uint8_t y) {
    return y == NUTDB_TT_RawStringLiteral ? 0 : (y == NUTDB_TT_EscapedSQStringLiteral ? 1 : 2);
  }

  // pops the top E_OP entry: BinaryOp{op, left, right}; refuses whatever simplify.rs would fold
  NUTDB_HD bool reduce() {
    const uint32_t e = einfo[--sp];
    const uint32_t op = (e >> 6) & 63u;
    const uint8_t lkind = (uint8_t)((e >> 12) & 255u);
    if (op == 9 || op == 10) {  // simplified_eq / simplified_neq
      if (is_literal(lkind) && is_literal(cur_kind)) return false;
    } else if (op >= 11 && op <= 13) {  // simplified_and / or / xor
      if (lkind == NUTDB_NK_LIT_BOOL || cur_kind == NUTDB_NK_LIT_BOOL) return false;
    }
    cur_start = emark[sp];
    cur_kind = NUTDB_NK_BINARY;
    return emit(NUTDB_NK_BINARY, (uint8_t)op, 0, cur_start);
  }

  // must_parse_expr (mod.rs:1205): on success the expression's subtree is [cur_start, n) and (ty, kw) is the
  // token that ended it.  There is exactly ONE call site (try_parse).
  NUTDB_HD bool expr() {
    for (;;) {
      // ---------------- operand: must_parse_expr_prefix (mod.rs:1222-1347) ----------------
      const uint32_t oi = F.opnd[ty];
      const uint32_t ocls = oi & 15u;
      if (ocls == FO_LEAF || ocls == FO_INT) {  // a literal (or `*`): one leaf, whatever its type
        if (ocls == FO_INT && !int_ok(ty, kw, 2)) return false;
        cur_start = n;
        cur_kind = (uint8_t)(oi >> 8);
        if (!emit((uint8_t)(oi >> 8), (uint8_t)(oi >> 16), (uint16_t)(oi >> 24), t)) return false;
        adv();
      } else if (ocls == FO_WORD || ocls == FO_DELIM) {
        if (ocls == FO_WORD) {
          if (kw == KW_TRUE || kw == KW_FALSE || kw == KW_NULL) {
            cur_start = n;
            cur_kind = kw == KW_NULL ? (uint8_t)NUTDB_NK_LIT_NULL : (uint8_t)NUTDB_NK_LIT_BOOL;
            if (!emit(cur_kind, kw == KW_TRUE ? 1 : 0, 0, NUTDB_CN_NOTOK)) return false;
            adv();
            goto operators;
          }
          if (kw == KW_NOT || kw == KW_INTERVAL || kw == KW_IF || kw == KW_CASE) return false;
        }
        const uint8_t ty2 = tok.type(t + 1);
        if (ty2 == NUTDB_TT_LParen && ocls == FO_WORD) {  // function call (mod.rs:1303-1308, :1538-1556)
          const uint8_t ty3 = tok.type(t + 2);
          if (ty3 == NUTDB_TT_KeywordOrIdentifier && (tok.kw(t + 2) == KW_SELECT || tok.kw(t + 2) == KW_WITH)) return false;
          const uint32_t m = n;
          if (!emit(NUTDB_NK_FN_NAME, 0, 0, t)) return false;
          if (ty3 == NUTDB_TT_RParen) {
            if (!emit(NUTDB_NK_FNCALL, 7, 0, m)) return false;
            cur_start = m;
            cur_kind = NUTDB_NK_FNCALL;
            adv(3);
          } else {
            if (sp >= DEPTH) return false;
            einfo[sp] = E_CALL;
            emark[sp] = m;
            sp++;
            adv(2);
            continue;  // first argument
          }
        } else if (ty2 == NUTDB_TT_Dot) {  // must_parse_identifier_based_prefix (mod.rs:1506-1523)
          const uint8_t ty3 = tok.type(t + 2);
          if (!(ident_string(ty3) || ty3 == NUTDB_TT_Mul)) return false;
          cur_start = n;
          cur_kind = NUTDB_NK_IDENT;
          if (!emit(NUTDB_NK_QUAL, 0, 0, t)) return false;
          if (!emit(NUTDB_NK_IDENT, ty3 == NUTDB_TT_Mul ? 1 : 0, 1, t + 2)) return false;
          adv(3);
        } else {
          cur_start = n;
          cur_kind = NUTDB_NK_IDENT;
          if (!emit(NUTDB_NK_IDENT, 0, 0, t)) return false;
          adv();
        }
      } else if (ocls == FO_ESTR) {
        if (!string_ok(t, ty)) return false;
        cur_start = n;
        cur_kind = NUTDB_NK_LIT_STR;
        if (!emit(NUTDB_NK_LIT_STR, (uint8_t)(oi >> 16), 0, t)) return false;
        adv();
      } else if (ocls == FO_MINUS) {  // only a literal may follow (mod.rs:1259-1269)
        const uint8_t ty2 = tok.type(t + 1);
        cur_start = n;
        if (ty2 == NUTDB_TT_FloatLiteral) {
          cur_kind = NUTDB_NK_LIT_FLOAT;
          if (!emit(NUTDB_NK_LIT_FLOAT, 1, 0, t + 1)) return false;
        } else if (ty2 == NUTDB_TT_IntegerLiteral || ty2 == NUTDB_TT_HexLiteral) {
          if (!int_ok(ty2, tok.kw(t + 1), 2)) return false;
          cur_kind = NUTDB_NK_LIT_INT;
          if (!emit(NUTDB_NK_LIT_INT, 1, ty2 == NUTDB_TT_HexLiteral ? 1 : 0, t + 1)) return false;
        } else {
          return false;
        }
        adv(2);
      } else if (ocls == FO_PLUS) {  // prefix plus is dropped (mod.rs:1270)
        adv();
        continue;
      } else if (ocls == FO_LPAREN) {  // (mod.rs:1229-1246); a subquery goes to the automaton
        if (next_is_kw(1, KW_SELECT) || next_is_kw(1, KW_WITH)) return false;
        if (sp >= DEPTH) return false;
        einfo[sp] = E_PAREN;
        emark[sp] = n;
        sp++;
        adv();
        continue;
      } else {
        return false;
      }
    operators:
      // ---------------- operators: token_power (mod.rs:1895-1927) + must_parse_expr_infix ----------------
      for (;;) {
        const uint32_t e = ty == NUTDB_TT_KeywordOrIdentifier ? F.opkw[kw] : F.optok[ty];
        if (e >> 12) return false;  // NOT / IS / BETWEEN / index access: the automaton
        const uint32_t power = e & 15u, op = (e >> 4) & 63u;
        // everything of equal or higher power on the stack is complete (left-associative)
        while (sp > 0 && (einfo[sp - 1] & 3u) == E_OP && ((einfo[sp - 1] >> 2) & 15u) >= power)
          if (!reduce()) return false;
        if (power != P_Terminator) {
          if (sp >= DEPTH) return false;
          einfo[sp] = E_OP | (power << 2) | (op << 6) | ((uint32_t)cur_kind << 12);
          emark[sp] = cur_start;
          sp++;
          adv();
          break;  // right operand
        }
        if (sp == 0) return true;  // the expression is complete
        // inside brackets opened by this expression
        const uint32_t btype = einfo[sp - 1] & 3u;
        if (ty == NUTDB_TT_Comma) {
          einfo[sp - 1] += 1u << 20;
          adv();
          break;  // next item
        }
        if (ty != NUTDB_TT_RParen) return false;
        const uint32_t items = (einfo[sp - 1] >> 20) + 1u;
        const uint32_t m = emark[sp - 1];
        sp--;
        adv();
        if (btype == E_CALL) {
          if (!emit(NUTDB_NK_FNCALL, 7, 0, m)) return false;
          cur_start = m;
          cur_kind = NUTDB_NK_FNCALL;
        } else if (items > 1) {  // one item in parentheses is the item itself (mod.rs:1236-1242)
          if (!emit(NUTDB_NK_COLLECTION, 0, 0, m)) return false;
          cur_start = m;
          cur_kind = NUTDB_NK_COLLECTION;
        }
      }
    }
  }

  NUTDB_HD bool alias() {  // [AS name] (mod.rs:563-578)
    if (is_kw(KW_AS)) {
      if (!ident_string(tok.type(t + 1))) return false;
      if (!emit(NUTDB_NK_ALIAS, 0, 0, t + 1)) return false;
      adv(2);
    }
    return true;
  }
  NUTDB_HD bool int_literal(uint32_t width_) {  // must_parse_integer_literal (mod.rs:1815) -> NK_NUM
    if (ty != NUTDB_TT_IntegerLiteral && ty != NUTDB_TT_HexLiteral) return false;
    if (!int_ok(ty, kw, width_)) return false;
    if (!emit(NUTDB_NK_NUM, 0, ty == NUTDB_TT_HexLiteral ? 1 : 0, t)) return false;
    adv();
    return true;
  }
  NUTDB_HD bool string_literal() {  // must_parse_string_literal (mod.rs:1833) -> NK_STR
    if (ty != NUTDB_TT_RawStringLiteral && ty != NUTDB_TT_EscapedSQStringLiteral && ty != NUTDB_TT_EscapedDQStringLiteral)
      return false;
    if (!string_ok(t, ty)) return false;
    if (!emit(NUTDB_NK_STR, str_sub(ty), 0, t)) return false;
    adv();
    return true;
  }

  // ---- SELECT (mod.rs:190-203, :279-544): what may follow once clause number `stage` is done ----
  // stages: 1 WHERE, 2 GROUP BY, 3 HAVING, 4 ORDER BY, 5 LIMIT, then the end of the body
  NUTDB_HD uint32_t select_advance(uint32_t stage) {
    if (stage <= 1 && is_kw(KW_WHERE)) {
      adv();
      m1 = n;
      ctx = C_WHERE;
      return R_EXPR;
    }
    if (stage <= 2 && is_kw(KW_GROUP)) {
      if (!next_is_kw(1, KW_BY)) return R_BAIL;
      adv(2);
      m1 = n;
      ctx = C_GROUP_ITEM;
      return R_EXPR;
    }
    if (stage <= 3 && is_kw(KW_HAVING)) {
      adv();
      m1 = n;
      ctx = C_HAVING;
      return R_EXPR;
    }
    if (stage <= 4 && is_kw(KW_ORDER)) {
      if (!next_is_kw(1, KW_BY)) return R_BAIL;
      adv(2);
      m1 = n;
      ctx = C_ORDER_ITEM;
      return R_EXPR;
    }
    if (is_kw(KW_LIMIT)) {  // mod.rs:503-544
      adv();
      const uint32_t m = n;
      uint32_t sub = 0, ax = 0;
      if (!int_literal(1)) return R_BAIL;
      if (ty == NUTDB_TT_Comma) {
        adv();
        sub = 1;
        if (!int_literal(1)) return R_BAIL;
      } else if (is_kw(KW_OFFSET)) {
        adv();
        sub = 2;
        if (!int_literal(1)) return R_BAIL;
      }
      if (is_kw(KW_WITH)) {
        if (!next_is_kw(1, KW_TIES)) return R_BAIL;
        adv(2);
        ax = 1;
      }
      if (!emit(NUTDB_NK_LIMIT, (uint8_t)sub, (uint16_t)ax, m)) return R_BAIL;
    }
    if (!emit(NUTDB_NK_QUERY_BODY, 0, 0, 0)) return R_BAIL;
    if (ty == NUTDB_TT_KeywordOrIdentifier && (kw == KW_UNION || kw == KW_INTERSECT || kw == KW_EXCEPT))
      return R_BAIL;  // set operations (mod.rs:250-267)
    return emit(NUTDB_NK_STMT_SELECT, 0, 0, 0) ? R_DONE : R_BAIL;
  }
  // the select list is complete: [FROM name [AS a]] (must_parse_query_source, mod.rs:546-569: a plain table name here)
  NUTDB_HD uint32_t select_after_items() {
    if (!emit(NUTDB_NK_COLS, 0, 0, 0)) return R_BAIL;
    if (is_kw(KW_FROM)) {
      adv();
      const uint32_t m = n;
      if (!ident_string(ty)) return R_BAIL;
      if (ty == NUTDB_TT_KeywordOrIdentifier) {
        if (kw == KW_TRUE || kw == KW_FALSE || kw == KW_NULL || kw == KW_NOT || kw == KW_INTERVAL || kw == KW_IF ||
            kw == KW_CASE)
          return R_BAIL;
        if (tok.type(t + 1) == NUTDB_TT_LParen) return R_BAIL;  // table function
      }
      if (tok.type(t + 1) == NUTDB_TT_Dot) return R_BAIL;       // qualified: the automaton drops the qualifier
      if (!emit(NUTDB_NK_IDENT, 0, 0, t)) return R_BAIL;
      adv();
      // the source is an expression: anything with infix power continues it (mod.rs:1212-1216)
      switch (ty) {
        case NUTDB_TT_Eq: case NUTDB_TT_NotEq: case NUTDB_TT_Gt: case NUTDB_TT_Lt: case NUTDB_TT_GtEq: case NUTDB_TT_LtEq:
        case NUTDB_TT_BitOr: case NUTDB_TT_BitXor: case NUTDB_TT_BitAnd: case NUTDB_TT_BitLShift: case NUTDB_TT_BitRShift:
        case NUTDB_TT_Plus: case NUTDB_TT_Minus: case NUTDB_TT_Mul: case NUTDB_TT_Div: case NUTDB_TT_Mod:
        case NUTDB_TT_LBracket:
          return R_BAIL;
        case NUTDB_TT_KeywordOrIdentifier:
          if (kw == KW_AND || kw == KW_OR || kw == KW_XOR || kw == KW_IN || kw == KW_LIKE || kw == KW_ILIKE ||
              kw == KW_NOT || kw == KW_IS || kw == KW_BETWEEN)
            return R_BAIL;
          break;
        default: break;
      }
      if (!alias()) return R_BAIL;
      if (!emit(NUTDB_NK_FROM, 0, 0, m)) return R_BAIL;
    }
    if (ty == NUTDB_TT_KeywordOrIdentifier &&
        (kw == KW_INNER || kw == KW_FULL || kw == KW_LEFT || kw == KW_RIGHT || kw == KW_JOIN))
      return R_BAIL;
    return select_advance(1);
  }

  // ---- CREATE TABLE (mod.rs:689-805, :936-972) ----
  // must_parse_datatype (mod.rs:1688-1797) without Enum / Tuple / Map
  NUTDB_HD bool datatype() {
    uint32_t marks[4], subs[4], depth = 0;
    for (;;) {
      if (ty != NUTDB_TT_KeywordOrIdentifier) return false;
      if (kw < KW_INT8 || kw > KW_NULLABLE) return false;
      const uint32_t i = kw - KW_INT8;
      adv();
      if (i == 26 || i == 30 || i == 31) {  // Array / Dictionary / Nullable (inner)
        if (depth >= 4 || ty != NUTDB_TT_LParen) return false;
        marks[depth] = n;
        subs[depth] = i == 26 ? 0u : (i == 30 ? 4u : 5u);
        depth++;
        adv();
        continue;
      }
      if (i > 25) return false;
      if (i == 16 || i == 17 || i == 21 || (i == 22 && ty == NUTDB_TT_LParen)) {
        if (ty != NUTDB_TT_LParen) return false;
        adv();
        const uint32_t m = n;
        if (!int_literal((i == 16 || i == 17) ? 0u : 1u)) return false;
        if (ty != NUTDB_TT_RParen) return false;
        adv();
        if (!emit(NUTDB_NK_DT_PARAM, (uint8_t)i, 0, m)) return false;
      } else {
        if (!emit(NUTDB_NK_DT_SCALAR, (uint8_t)i, 0, NUTDB_CN_NOTOK)) return false;
      }
      break;
    }
    while (depth > 0) {
      depth--;
      if (ty != NUTDB_TT_RParen) return false;
      adv();
      if (!emit(NUTDB_NK_DT_COMPOUND, (uint8_t)subs[depth], 0, marks[depth])) return false;
    }
    return true;
  }
  // table attributes after the column list (mod.rs:746-803)
  NUTDB_HD uint32_t table_attrs() {
    while (ty == NUTDB_TT_KeywordOrIdentifier) {
      uint32_t bit;
      uint8_t next_ctx;
      if (kw == KW_PRIMARY) { bit = 1; next_ctx = C_TBL_PK_ITEM; }
      else if (kw == KW_ORDER) { bit = 2; next_ctx = C_TBL_ORDER_ITEM; }
      else if (kw == KW_PARTITION) { bit = 4; next_ctx = C_TBL_PART; }
      else if (kw == KW_COMMENT) { bit = 8; next_ctx = 0xFF; }
      else return R_BAIL;
      if (tseen & bit) return R_BAIL;  // Conflicts: the automaton reports it
      tseen |= bit;
      adv();
      if (next_ctx == 0xFF) {
        if (!string_literal()) return R_BAIL;
        continue;
      }
      if (!is_kw(next_ctx == C_TBL_PK_ITEM ? (uint32_t)KW_KEY : (uint32_t)KW_BY)) return R_BAIL;
      adv();
      m1 = n;
      ctx = next_ctx;
      return R_EXPR;
    }
    if (!emit(NUTDB_NK_TABLEDEF, 0, 0, 0)) return R_BAIL;
    return emit(NUTDB_NK_STMT_CREATE, 0, (uint16_t)aux, 0) ? R_DONE : R_BAIL;
  }
  // attributes of the column in progress, then the next column or the end of the list (mod.rs:936-972, :722-733)
  NUTDB_HD uint32_t column_attrs() {
    for (;;) {
      while (ty == NUTDB_TT_KeywordOrIdentifier) {
        if (kw == KW_DEFAULT) {
          if (seen & 1u) return R_BAIL;
          seen |= 1u;
          adv();
          m1 = n;
          ctx = C_COL_DEFAULT;
          return R_EXPR;
        }
        if (kw != KW_COMMENT || (seen & 2u)) return R_BAIL;
        seen |= 2u;
        adv();
        if (!string_literal()) return R_BAIL;
      }
      if (!emit(NUTDB_NK_COLDEF, 0, 0, m0)) return R_BAIL;
      if (ty != NUTDB_TT_Comma) break;
      adv();
      if (!begin_column()) return R_BAIL;
    }
    if (ty != NUTDB_TT_RParen) return R_BAIL;
    adv();
    tseen = 0;
    return table_attrs();
  }
  NUTDB_HD bool begin_column() {
    if (is_kw(KW_INDEX) || is_kw(KW_CONSTRAINT)) return false;
    m0 = n;
    seen = 0;
    if (!ident_string(ty)) return false;
    if (!emit(NUTDB_NK_NAME, 0, 0, t)) return false;
    adv();
    return datatype();
  }

  // ---- statement prologues: everything before the first expression ----
  NUTDB_HD uint32_t begin_select() {
    if (is_kw(KW_DISTINCT)) return R_BAIL;
    ctx = C_SEL_ITEM;
    return R_EXPR;
  }
  NUTDB_HD uint32_t begin_insert() {  // try_parse_insert_stmt with VALUES (mod.rs:589-670)
    if (!is_kw(KW_INTO)) return R_BAIL;
    adv();
    if (!ident_string(ty)) return R_BAIL;
    if (!emit(NUTDB_NK_NAME, 0, 0, t)) return R_BAIL;
    adv();
    if (ty == NUTDB_TT_LParen) {
      adv();
      for (;;) {
        if (!ident_string(ty)) return R_BAIL;
        if (!emit(NUTDB_NK_NAME, 0, 0, t)) return R_BAIL;
        adv();
        if (ty != NUTDB_TT_Comma) break;
        adv();
      }
      if (ty != NUTDB_TT_RParen) return R_BAIL;
      adv();
    }
    if (!is_kw(KW_VALUES)) return R_BAIL;
    adv();
    m0 = n;
    if (ty != NUTDB_TT_LParen) return R_BAIL;
    adv();
    m1 = n;
    w = 0;
    row = 0;
    ctx = C_INS_VALUE;
    return R_EXPR;
  }
  NUTDB_HD uint32_t begin_create() {
    if (!is_kw(KW_TABLE)) return R_BAIL;
    adv();
    aux = 0;
    if (is_kw(KW_IF)) {
      if (!next_is_kw(1, KW_NOT) || !next_is_kw(2, KW_EXISTS)) return R_BAIL;
      adv(3);
      aux = 1;
    }
    if (!ident_string(ty)) return R_BAIL;
    if (!emit(NUTDB_NK_NAME, 0, 0, t)) return R_BAIL;
    adv();
    if (ty != NUTDB_TT_LParen) return R_BAIL;
    adv();
    if (!begin_column()) return R_BAIL;
    return column_attrs();
  }

  // what follows the expression that was just parsed in context `ctx`
  NUTDB_HD uint32_t after_expr() {
    switch (ctx) {
      case C_SEL_ITEM:
        if (!alias()) return R_BAIL;
        if (ty == NUTDB_TT_Comma) {
          adv();
          return R_EXPR;
        }
        return select_after_items();
      case C_WHERE:
        if (!emit(NUTDB_NK_WHERE, 0, 0, m1)) return R_BAIL;
        return select_advance(2);
      case C_GROUP_ITEM:
        if (!alias()) return R_BAIL;
        if (ty == NUTDB_TT_Comma) {
          adv();
          return R_EXPR;
        }
        if (!emit(NUTDB_NK_GROUPBY, 0, 0, m1)) return R_BAIL;
        return select_advance(3);
      case C_HAVING:
        if (!emit(NUTDB_NK_HAVING, 0, 0, m1)) return R_BAIL;
        return select_advance(4);
      case C_ORDER_ITEM:  // DESC only: the reference never accepts ASC (mod.rs:491-496)
        if (!alias()) return R_BAIL;
        if (is_kw(KW_DESC)) {
          if (!emit(NUTDB_NK_ORDER_DESC, 0, 0, NUTDB_CN_NOTOK)) return R_BAIL;
          adv();
        }
        if (ty == NUTDB_TT_Comma) {
          adv();
          return R_EXPR;
        }
        if (!emit(NUTDB_NK_ORDERBY, 0, 0, m1)) return R_BAIL;
        return select_advance(5);
      case C_INS_VALUE:  // must_parse_insert_rows (mod.rs:636-670)
        w++;
        if (ty == NUTDB_TT_Comma) {
          adv();
          return R_EXPR;
        }
        if (!emit(NUTDB_NK_ROW, 0, 0, m1)) return R_BAIL;
        if (row == 0) width = w;
        else if (w != width) return R_BAIL;  // Conflicts: the automaton reports it
        if (ty != NUTDB_TT_RParen) return R_BAIL;
        adv();
        if (ty == NUTDB_TT_Comma) {
          adv();
          if (ty != NUTDB_TT_LParen) return R_BAIL;
          adv();
          m1 = n;
          w = 0;
          row++;
          return R_EXPR;
        }
        if (!emit(NUTDB_NK_ROWS, 0, 0, m0)) return R_BAIL;
        return emit(NUTDB_NK_STMT_INSERT, 0, 0, 0) ? R_DONE : R_BAIL;
      case C_COL_DEFAULT:
        if (!emit(NUTDB_NK_ATTR_DEFAULT, 0, 0, m1)) return R_BAIL;
        return column_attrs();
      case C_TBL_PK_ITEM:
      case C_TBL_ORDER_ITEM:
        if (ty == NUTDB_TT_Comma) {
          adv();
          return R_EXPR;
        }
        if (!emit(ctx == C_TBL_PK_ITEM ? (uint8_t)NUTDB_NK_ATTR_PK : (uint8_t)NUTDB_NK_ATTR_ORDER, 0, 0, m1)) return R_BAIL;
        return table_attrs();
      default:  // C_TBL_PART
        if (!emit(NUTDB_NK_ATTR_PART, 0, 0, m1)) return R_BAIL;
        return table_attrs();
    }
  }

  // parse_stmt (mod.rs:128-180).  true: res describes a successful parse with n nodes emitted.
  NUTDB_HD bool try_parse(ParseResult& res) {
    t = 0;
    load();
    if (ty != NUTDB_TT_KeywordOrIdentifier) return false;
    const uint32_t first = kw;
    adv();
    uint32_t r;
    if (first == KW_SELECT) r = begin_select();
    else if (first == KW_INSERT) r = begin_insert();
    else if (first == KW_CREATE) r = begin_create();
    else return false;
    while (r == R_EXPR) {
      if (!expr()) return false;
      r = after_expr();
    }
    if (r != R_DONE || sp != 0) return false;
    if (ty != NUTDB_TT_EOF && ty != NUTDB_TT_SemiColon) return false;
    res.status = NUTDB_ST_OK;
    res.node_count = n;
    res.tok_used = t + 1;
    res.err_code = 0;
    res.err_has_pos = false;
    res.err_pos = res.err_a = res.err_b = res.err_c = 0;
    return true;
  }
};

}  // namespace npar
