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

template <class Tok, class Nodes, class Text>
struct FastParser {
  Tok& tok;
  Nodes& nd;
  Text& text;
  uint32_t t = 0, n = 0, cap;
  // the operand being built (right-most subtree)
  uint32_t cur_start = 0;
  uint8_t cur_kind = 0;
  // operator / bracket stack
  enum : uint32_t { E_OP = 0, E_PAREN = 1, E_CALL = 2 };
  static const uint32_t DEPTH = 12;
  uint32_t einfo[DEPTH];   // type | power << 2 | op << 6 | left kind << 12 | item count << 20
  uint32_t emark[DEPTH];   // E_OP: start of the left operand; brackets: node count at the opening
  uint32_t sp = 0;

  NUTDB_HD FastParser(Tok& tk, Nodes& nodes, Text& tx) : tok(tk), nd(nodes), text(tx), cap(nodes.capacity()) {}

  NUTDB_HD bool emit(uint8_t kind, uint8_t sub, uint16_t aux, uint32_t x) {
    if (n >= cap) return false;
    CNode c;
    c.kind = kind;
    c.sub = sub;
    c.aux = aux;
    c.x = x;
    nd.set(n++, c);
    return true;
  }
  NUTDB_HD static bool is_literal(uint8_t kind) { return kind >= NUTDB_NK_LIT_INT && kind <= NUTDB_NK_LIT_INTERVAL; }
  NUTDB_HD bool is_kw(uint32_t i, uint32_t kw) { return tok.type(i) == NUTDB_TT_KeywordOrIdentifier && tok.kw(i) == kw; }
  NUTDB_HD static bool ident_string(uint8_t ty) {  // must_parse_identifier_string (mod.rs:1682)
    return ty == NUTDB_TT_KeywordOrIdentifier || ty == NUTDB_TT_DelimitedIdentifier;
  }
  // integer_from_str! cannot fail for 1..safe digits; the lexer stores min(len, 255) in the kw byte
  NUTDB_HD bool int_ok(uint32_t i, uint32_t width) {
    const bool hex = tok.type(i) == NUTDB_TT_HexLiteral;
    const uint32_t len = tok.kw(i);
    const uint32_t safe = width == 0 ? 2u : width == 1 ? (hex ? 16u : 19u) : (hex ? 32u : 38u);
    return len >= 1 && len <= safe;
  }
  // an escaped string literal can only be rejected through a backslash-u escape (literal.rs:70-88)
  NUTDB_HD bool string_ok(uint32_t i) {
    if (tok.type(i) == NUTDB_TT_RawStringLiteral) return true;
    const uint32_t s = tok.start(i), e = tok.end(i);
    for (uint32_t p = s; p + 1 < e; p++)
      if (text.byte(p) == '\\' && text.byte(p + 1) == 'u') return false;
    return true;
  }
  NUTDB_HD static uint8_t str_sub(uint8_t ty) {
    return ty == NUTDB_TT_RawStringLiteral ? 0 : (ty == NUTDB_TT_EscapedSQStringLiteral ? 1 : 2);
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

  // must_parse_expr (mod.rs:1205): on success the expression's subtree is [cur_start, n)
  NUTDB_HD bool expr() {
    const uint32_t base = sp;
    for (;;) {
      // ---------------- operand: must_parse_expr_prefix (mod.rs:1222-1347) ----------------
      const uint8_t ty = tok.type(t);
      switch (ty) {
        case NUTDB_TT_KeywordOrIdentifier: {
          const uint32_t kw = tok.kw(t);
          if (kw == KW_TRUE || kw == KW_FALSE) {
            cur_start = n;
            cur_kind = NUTDB_NK_LIT_BOOL;
            if (!emit(NUTDB_NK_LIT_BOOL, kw == KW_TRUE ? 1 : 0, 0, NUTDB_CN_NOTOK)) return false;
            t++;
            break;
          }
          if (kw == KW_NULL) {
            cur_start = n;
            cur_kind = NUTDB_NK_LIT_NULL;
            if (!emit(NUTDB_NK_LIT_NULL, 0, 0, NUTDB_CN_NOTOK)) return false;
            t++;
            break;
          }
          if (kw == KW_NOT || kw == KW_INTERVAL || kw == KW_IF || kw == KW_CASE) return false;
          const uint8_t ty2 = tok.type(t + 1);
          if (ty2 == NUTDB_TT_LParen) {  // function call (mod.rs:1303-1308, :1538-1556)
            const uint8_t ty3 = tok.type(t + 2);
            if (ty3 == NUTDB_TT_KeywordOrIdentifier && (tok.kw(t + 2) == KW_SELECT || tok.kw(t + 2) == KW_WITH))
              return false;
            const uint32_t m = n;
            if (!emit(NUTDB_NK_FN_NAME, 0, 0, t)) return false;
            if (ty3 == NUTDB_TT_RParen) {
              if (!emit(NUTDB_NK_FNCALL, 7, 0, m)) return false;
              cur_start = m;
              cur_kind = NUTDB_NK_FNCALL;
              t += 3;
              break;
            }
            if (sp >= DEPTH) return false;
            einfo[sp] = E_CALL;
            emark[sp] = m;
            sp++;
            t += 2;
            continue;  // first argument
          }
          // fallthrough to the identifier forms
        }
        case NUTDB_TT_DelimitedIdentifier: {  // must_parse_identifier_based_prefix (mod.rs:1506-1523)
          if (tok.type(t + 1) == NUTDB_TT_Dot) {
            const uint8_t ty3 = tok.type(t + 2);
            if (!(ident_string(ty3) || ty3 == NUTDB_TT_Mul)) return false;
            cur_start = n;
            cur_kind = NUTDB_NK_IDENT;
            if (!emit(NUTDB_NK_QUAL, 0, 0, t)) return false;
            if (!emit(NUTDB_NK_IDENT, ty3 == NUTDB_TT_Mul ? 1 : 0, 1, t + 2)) return false;
            t += 3;
          } else {
            cur_start = n;
            cur_kind = NUTDB_NK_IDENT;
            if (!emit(NUTDB_NK_IDENT, 0, 0, t)) return false;
            t++;
          }
          break;
        }
        case NUTDB_TT_Mul:
          cur_start = n;
          cur_kind = NUTDB_NK_IDENT;
          if (!emit(NUTDB_NK_IDENT, 1, 0, t)) return false;
          t++;
          break;
        case NUTDB_TT_RawStringLiteral:
        case NUTDB_TT_EscapedSQStringLiteral:
        case NUTDB_TT_EscapedDQStringLiteral:
          if (!string_ok(t)) return false;
          cur_start = n;
          cur_kind = NUTDB_NK_LIT_STR;
          if (!emit(NUTDB_NK_LIT_STR, str_sub(ty), 0, t)) return false;
          t++;
          break;
        case NUTDB_TT_FloatLiteral:
          cur_start = n;
          cur_kind = NUTDB_NK_LIT_FLOAT;
          if (!emit(NUTDB_NK_LIT_FLOAT, 0, 0, t)) return false;
          t++;
          break;
        case NUTDB_TT_IntegerLiteral:
        case NUTDB_TT_HexLiteral:
          if (!int_ok(t, 2)) return false;
          cur_start = n;
          cur_kind = NUTDB_NK_LIT_INT;
          if (!emit(NUTDB_NK_LIT_INT, 0, ty == NUTDB_TT_HexLiteral ? 1 : 0, t)) return false;
          t++;
          break;
        case NUTDB_TT_Minus: {  // only a literal may follow (mod.rs:1259-1269)
          const uint8_t ty2 = tok.type(t + 1);
          cur_start = n;
          if (ty2 == NUTDB_TT_FloatLiteral) {
            cur_kind = NUTDB_NK_LIT_FLOAT;
            if (!emit(NUTDB_NK_LIT_FLOAT, 1, 0, t + 1)) return false;
          } else if (ty2 == NUTDB_TT_IntegerLiteral || ty2 == NUTDB_TT_HexLiteral) {
            if (!int_ok(t + 1, 2)) return false;
            cur_kind = NUTDB_NK_LIT_INT;
            if (!emit(NUTDB_NK_LIT_INT, 1, ty2 == NUTDB_TT_HexLiteral ? 1 : 0, t + 1)) return false;
          } else {
            return false;
          }
          t += 2;
          break;
        }
        case NUTDB_TT_Plus:  // prefix plus is dropped (mod.rs:1270)
          t++;
          continue;
        case NUTDB_TT_LParen: {  // (mod.rs:1229-1246); a subquery goes to the automaton
          if (tok.type(t + 1) == NUTDB_TT_KeywordOrIdentifier && (tok.kw(t + 1) == KW_SELECT || tok.kw(t + 1) == KW_WITH))
            return false;
          if (sp >= DEPTH) return false;
          einfo[sp] = E_PAREN;
          emark[sp] = n;
          sp++;
          t++;
          continue;
        }
        default: return false;
      }
      // ---------------- operators: token_power (mod.rs:1895-1927) + must_parse_expr_infix ----------------
      for (;;) {
        const uint8_t oy = tok.type(t);
        uint32_t power = P_Terminator, op = 0;
        switch (oy) {
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
          case NUTDB_TT_LBracket: return false;  // index access
          case NUTDB_TT_KeywordOrIdentifier:
            switch (tok.kw(t)) {
              case KW_AND: power = P_And; op = 11; break;
              case KW_OR: power = P_Or; op = 12; break;
              case KW_XOR: power = P_Xor; op = 13; break;
              case KW_IN: power = P_Comparison; op = 18; break;
              case KW_LIKE: power = P_Comparison; op = 14; break;
              case KW_ILIKE: power = P_Comparison; op = 16; break;
              case KW_NOT: case KW_IS: case KW_BETWEEN: return false;
              default: break;
            }
            break;
          default: break;
        }
        // everything of equal or higher power on the stack is complete (left-associative)
        while (sp > base && (einfo[sp - 1] & 3u) == E_OP && ((einfo[sp - 1] >> 2) & 15u) >= power)
          if (!reduce()) return false;
        if (power != P_Terminator) {
          if (sp >= DEPTH) return false;
          einfo[sp] = E_OP | (power << 2) | (op << 6) | ((uint32_t)cur_kind << 12);
          emark[sp] = cur_start;
          sp++;
          t++;
          break;  // right operand
        }
        if (sp == base) return true;  // the expression is complete
        // inside brackets opened by this expression
        const uint32_t btype = einfo[sp - 1] & 3u;
        if (oy == NUTDB_TT_Comma) {
          einfo[sp - 1] += 1u << 20;
          t++;
          break;  // next item
        }
        if (oy != NUTDB_TT_RParen) return false;
        const uint32_t items = (einfo[sp - 1] >> 20) + 1u;
        const uint32_t m = emark[sp - 1];
        sp--;
        t++;
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

  // must_parse_query_expr (mod.rs:571-579): expr [AS name]
  NUTDB_HD bool query_expr() {
    if (!expr()) return false;
    if (is_kw(t, KW_AS)) {
      if (!ident_string(tok.type(t + 1))) return false;
      if (!emit(NUTDB_NK_ALIAS, 0, 0, t + 1)) return false;
      t += 2;
    }
    return true;
  }
  NUTDB_HD bool query_expr_list() {
    for (;;) {
      if (!query_expr()) return false;
      if (tok.type(t) != NUTDB_TT_Comma) return true;
      t++;
    }
  }
  NUTDB_HD bool expr_list() {
    for (;;) {
      if (!expr()) return false;
      if (tok.type(t) != NUTDB_TT_Comma) return true;
      t++;
    }
  }
  NUTDB_HD bool int_literal(uint32_t width) {  // must_parse_integer_literal (mod.rs:1815) -> NK_NUM
    const uint8_t ty = tok.type(t);
    if (ty != NUTDB_TT_IntegerLiteral && ty != NUTDB_TT_HexLiteral) return false;
    if (!int_ok(t, width)) return false;
    if (!emit(NUTDB_NK_NUM, 0, ty == NUTDB_TT_HexLiteral ? 1 : 0, t)) return false;
    t++;
    return true;
  }
  NUTDB_HD bool string_literal() {  // must_parse_string_literal (mod.rs:1833) -> NK_STR
    const uint8_t ty = tok.type(t);
    if (ty != NUTDB_TT_RawStringLiteral && ty != NUTDB_TT_EscapedSQStringLiteral && ty != NUTDB_TT_EscapedDQStringLiteral)
      return false;
    if (!string_ok(t)) return false;
    if (!emit(NUTDB_NK_STR, str_sub(ty), 0, t)) return false;
    t++;
    return true;
  }

  // try_parse_select_stmt / must_parse_query_body (mod.rs:190-203, :279-325); t is after SELECT
  NUTDB_HD bool select_stmt() {
    if (is_kw(t, KW_DISTINCT)) return false;
    const uint32_t body = n;
    if (!query_expr_list()) return false;
    if (!emit(NUTDB_NK_COLS, 0, 0, body)) return false;
    if (is_kw(t, KW_FROM)) {  // must_parse_query_source (mod.rs:546-569): a plain table name here
      t++;
      const uint32_t m = n;
      const uint8_t ty = tok.type(t);
      if (!ident_string(ty)) return false;
      if (ty == NUTDB_TT_KeywordOrIdentifier) {
        const uint32_t kw = tok.kw(t);
        if (kw == KW_TRUE || kw == KW_FALSE || kw == KW_NULL || kw == KW_NOT || kw == KW_INTERVAL || kw == KW_IF ||
            kw == KW_CASE)
          return false;
      }
      if (!expr()) return false;
      if (cur_kind != NUTDB_NK_IDENT || cur_start != m || n != m + 1) return false;  // only `name`
      if (is_kw(t, KW_AS)) {
        if (!ident_string(tok.type(t + 1))) return false;
        if (!emit(NUTDB_NK_ALIAS, 0, 0, t + 1)) return false;
        t += 2;
      }
      if (!emit(NUTDB_NK_FROM, 0, 0, m)) return false;
    }
    if (tok.type(t) == NUTDB_TT_KeywordOrIdentifier) {
      const uint32_t kw = tok.kw(t);
      if (kw == KW_INNER || kw == KW_FULL || kw == KW_LEFT || kw == KW_RIGHT || kw == KW_JOIN) return false;
    }
    if (is_kw(t, KW_WHERE)) {
      t++;
      const uint32_t m = n;
      if (!expr()) return false;
      if (!emit(NUTDB_NK_WHERE, 0, 0, m)) return false;
    }
    if (is_kw(t, KW_GROUP)) {
      if (!is_kw(t + 1, KW_BY)) return false;
      t += 2;
      const uint32_t m = n;
      if (!query_expr_list()) return false;
      if (!emit(NUTDB_NK_GROUPBY, 0, 0, m)) return false;
    }
    if (is_kw(t, KW_HAVING)) {
      t++;
      const uint32_t m = n;
      if (!expr()) return false;
      if (!emit(NUTDB_NK_HAVING, 0, 0, m)) return false;
    }
    if (is_kw(t, KW_ORDER)) {  // DESC only: the reference never accepts ASC (mod.rs:491-496)
      if (!is_kw(t + 1, KW_BY)) return false;
      t += 2;
      const uint32_t m = n;
      for (;;) {
        if (!query_expr()) return false;
        if (is_kw(t, KW_DESC)) {
          if (!emit(NUTDB_NK_ORDER_DESC, 0, 0, NUTDB_CN_NOTOK)) return false;
          t++;
        }
        if (tok.type(t) != NUTDB_TT_Comma) break;
        t++;
      }
      if (!emit(NUTDB_NK_ORDERBY, 0, 0, m)) return false;
    }
    if (is_kw(t, KW_LIMIT)) {  // mod.rs:503-544
      t++;
      const uint32_t m = n;
      uint32_t sub = 0, aux = 0;
      if (!int_literal(1)) return false;
      if (tok.type(t) == NUTDB_TT_Comma) {
        t++;
        sub = 1;
        if (!int_literal(1)) return false;
      } else if (is_kw(t, KW_OFFSET)) {
        t++;
        sub = 2;
        if (!int_literal(1)) return false;
      }
      if (is_kw(t, KW_WITH)) {
        if (!is_kw(t + 1, KW_TIES)) return false;
        t += 2;
        aux = 1;
      }
      if (!emit(NUTDB_NK_LIMIT, (uint8_t)sub, (uint16_t)aux, m)) return false;
    }
    if (!emit(NUTDB_NK_QUERY_BODY, 0, 0, body)) return false;
    if (tok.type(t) == NUTDB_TT_KeywordOrIdentifier) {  // set operations (mod.rs:250-267)
      const uint32_t kw = tok.kw(t);
      if (kw == KW_UNION || kw == KW_INTERSECT || kw == KW_EXCEPT) return false;
    }
    return emit(NUTDB_NK_STMT_SELECT, 0, 0, 0);
  }

  // try_parse_insert_stmt with VALUES (mod.rs:589-670); t is after INSERT
  NUTDB_HD bool insert_stmt() {
    if (!is_kw(t, KW_INTO)) return false;
    t++;
    if (!ident_string(tok.type(t))) return false;
    if (!emit(NUTDB_NK_NAME, 0, 0, t)) return false;
    t++;
    if (tok.type(t) == NUTDB_TT_LParen) {
      t++;
      for (;;) {
        if (!ident_string(tok.type(t))) return false;
        if (!emit(NUTDB_NK_NAME, 0, 0, t)) return false;
        t++;
        if (tok.type(t) != NUTDB_TT_Comma) break;
        t++;
      }
      if (tok.type(t) != NUTDB_TT_RParen) return false;
      t++;
    }
    if (!is_kw(t, KW_VALUES)) return false;
    t++;
    const uint32_t rows = n;
    uint32_t width = 0;
    for (uint32_t r = 0;; r++) {
      if (tok.type(t) != NUTDB_TT_LParen) return false;
      t++;
      const uint32_t m = n;
      uint32_t w = 0;
      for (;;) {
        if (!expr()) return false;
        w++;
        if (tok.type(t) != NUTDB_TT_Comma) break;
        t++;
      }
      if (!emit(NUTDB_NK_ROW, 0, 0, m)) return false;
      if (r == 0) width = w;
      else if (w != width) return false;  // Conflicts: the automaton reports it
      if (tok.type(t) != NUTDB_TT_RParen) return false;
      t++;
      if (tok.type(t) != NUTDB_TT_Comma) break;
      t++;
    }
    if (!emit(NUTDB_NK_ROWS, 0, 0, rows)) return false;
    return emit(NUTDB_NK_STMT_INSERT, 0, 0, 0);
  }

  // must_parse_datatype (mod.rs:1688-1797) without Enum / Tuple / Map
  NUTDB_HD bool datatype() {
    uint32_t marks[4], subs[4], depth = 0;
    for (;;) {
      if (tok.type(t) != NUTDB_TT_KeywordOrIdentifier) return false;
      const uint32_t kw = tok.kw(t);
      if (kw < KW_INT8 || kw > KW_NULLABLE) return false;
      const uint32_t i = kw - KW_INT8;
      t++;
      if (i == 26 || i == 30 || i == 31) {  // Array / Dictionary / Nullable (inner)
        if (depth >= 4 || tok.type(t) != NUTDB_TT_LParen) return false;
        marks[depth] = n;
        subs[depth] = i == 26 ? 0u : (i == 30 ? 4u : 5u);
        depth++;
        t++;
        continue;
      }
      if (i > 25) return false;
      if (i == 16 || i == 17 || i == 21 || (i == 22 && tok.type(t) == NUTDB_TT_LParen)) {
        if (tok.type(t) != NUTDB_TT_LParen) return false;
        t++;
        const uint32_t m = n;
        if (!int_literal((i == 16 || i == 17) ? 0u : 1u)) return false;
        if (tok.type(t) != NUTDB_TT_RParen) return false;
        t++;
        if (!emit(NUTDB_NK_DT_PARAM, (uint8_t)i, 0, m)) return false;
      } else {
        if (!emit(NUTDB_NK_DT_SCALAR, (uint8_t)i, 0, NUTDB_CN_NOTOK)) return false;
      }
      break;
    }
    while (depth > 0) {
      depth--;
      if (tok.type(t) != NUTDB_TT_RParen) return false;
      t++;
      if (!emit(NUTDB_NK_DT_COMPOUND, (uint8_t)subs[depth], 0, marks[depth])) return false;
    }
    return true;
  }

  // try_parse_create_stmt for tables (mod.rs:689-805, :936-972); t is after CREATE
  NUTDB_HD bool create_stmt() {
    if (!is_kw(t, KW_TABLE)) return false;
    t++;
    uint32_t aux = 0;
    if (is_kw(t, KW_IF)) {
      if (!is_kw(t + 1, KW_NOT) || !is_kw(t + 2, KW_EXISTS)) return false;
      t += 3;
      aux = 1;
    }
    if (!ident_string(tok.type(t))) return false;
    if (!emit(NUTDB_NK_NAME, 0, 0, t)) return false;
    t++;
    if (tok.type(t) != NUTDB_TT_LParen) return false;
    t++;
    for (;;) {
      if (is_kw(t, KW_INDEX) || is_kw(t, KW_CONSTRAINT)) return false;
      const uint32_t m = n;
      if (!ident_string(tok.type(t))) return false;
      if (!emit(NUTDB_NK_NAME, 0, 0, t)) return false;
      t++;
      if (!datatype()) return false;
      uint32_t seen = 0;
      while (tok.type(t) == NUTDB_TT_KeywordOrIdentifier) {
        const uint32_t kw = tok.kw(t);
        if (kw == KW_DEFAULT) {
          if (seen & 1u) return false;
          seen |= 1u;
          t++;
          const uint32_t d = n;
          if (!expr()) return false;
          if (!emit(NUTDB_NK_ATTR_DEFAULT, 0, 0, d)) return false;
        } else if (kw == KW_COMMENT) {
          if (seen & 2u) return false;
          seen |= 2u;
          t++;
          if (!string_literal()) return false;
        } else {
          return false;
        }
      }
      if (!emit(NUTDB_NK_COLDEF, 0, 0, m)) return false;
      if (tok.type(t) != NUTDB_TT_Comma) break;
      t++;
    }
    if (tok.type(t) != NUTDB_TT_RParen) return false;
    t++;
    uint32_t seen = 0;
    while (tok.type(t) == NUTDB_TT_KeywordOrIdentifier) {
      const uint32_t kw = tok.kw(t);
      uint32_t bit, kind;
      if (kw == KW_PRIMARY) { bit = 1; kind = NUTDB_NK_ATTR_PK; }
      else if (kw == KW_ORDER) { bit = 2; kind = NUTDB_NK_ATTR_ORDER; }
      else if (kw == KW_PARTITION) { bit = 4; kind = NUTDB_NK_ATTR_PART; }
      else if (kw == KW_COMMENT) { bit = 8; kind = 0; }
      else return false;
      if (seen & bit) return false;
      seen |= bit;
      t++;
      if (kind == 0) {
        if (!string_literal()) return false;
        continue;
      }
      if (!is_kw(t, kind == NUTDB_NK_ATTR_PK ? (uint32_t)KW_KEY : (uint32_t)KW_BY)) return false;
      t++;
      const uint32_t m = n;
      if (kind == NUTDB_NK_ATTR_PART) {
        if (!expr()) return false;
      } else {
        if (!expr_list()) return false;
      }
      if (!emit((uint8_t)kind, 0, 0, m)) return false;
    }
    if (!emit(NUTDB_NK_TABLEDEF, 0, 0, 0)) return false;
    return emit(NUTDB_NK_STMT_CREATE, 0, (uint16_t)aux, 0);
  }

  // parse_stmt (mod.rs:128-180).  true: res describes a successful parse with n nodes emitted.
  NUTDB_HD bool try_parse(ParseResult& res) {
    if (tok.type(0) != NUTDB_TT_KeywordOrIdentifier) return false;
    const uint32_t kw = tok.kw(0);
    t = 1;
    bool ok;
    if (kw == KW_SELECT) ok = select_stmt();
    else if (kw == KW_INSERT) ok = insert_stmt();
    else if (kw == KW_CREATE) ok = create_stmt();
    else return false;
    if (!ok || sp != 0) return false;
    const uint8_t ty = tok.type(t);
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
