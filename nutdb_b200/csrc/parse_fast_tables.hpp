// parse_fast_tables.hpp -- builds the transition table of the table-driven parser in parse_fast.cuh (host only).
//
// Each row is one position of the grammar (reference src/parser/mod.rs; line numbers in the comments), each
// column a token class; an entry names a transition record: what to check, which node to emit, whether the token
// is consumed, which marks / context to set and the next state.  Anything not listed bails to the exact automaton.
#pragma once
#include <cstring>
#include <map>
#include <stdexcept>
#include <utility>
#include <vector>

#include "parse_fast.cuh"

namespace npar {

struct FastRec {
  uint32_t lo = FA_STEP, hi = 0;
  FastRec& act(uint32_t a) { lo = (lo & ~15u) | a; return *this; }
  FastRec& to(uint32_t s) { lo = (lo & ~(127u << 4)) | (s << 4); return *this; }
  FastRec& adv() { lo |= 1u << 11; return *this; }
  FastRec& emit(uint32_t mode, uint32_t kind, uint32_t sub = 0, uint32_t auxbit = 0) {
    lo |= (mode << 12) | (kind << 15) | (sub << 23) | (auxbit << 27);
    return *this;
  }
  FastRec& leaf(uint32_t kind, uint32_t sub = 0, uint32_t auxbit = 0) { return emit(FE_LEAF_TOK, kind, sub, auxbit); }
  FastRec& leaf0(uint32_t kind, uint32_t sub = 0) { return emit(FE_LEAF_NOTOK, kind, sub); }
  FastRec& node0(uint32_t kind, uint32_t sub = 0, uint32_t auxbit = 0) { return emit(FE_NODE_ZERO, kind, sub, auxbit); }
  FastRec& node_m0(uint32_t kind) { return emit(FE_NODE_M0, kind); }
  FastRec& node_m1(uint32_t kind, uint32_t sub = 0, uint32_t auxbit = 0) { return emit(FE_NODE_M1, kind, sub, auxbit); }
  FastRec& auxreg() { lo |= 1u << 28; return *this; }
  FastRec& pre_m0() { hi |= 1u; return *this; }
  FastRec& pre_m1() { hi |= 2u; return *this; }
  FastRec& post_m0() { hi |= 4u; return *this; }
  FastRec& post_m1() { hi |= 8u; return *this; }
  FastRec& cur() { hi |= 16u; return *this; }
  FastRec& check(uint32_t c) { hi |= c << 5; return *this; }
  FastRec& look(uint32_t l) { hi |= l << 8; return *this; }
  FastRec& ctx(uint32_t c) { hi |= (1u << 10) | (c << 11); return *this; }
  FastRec& bit(uint32_t b) { hi |= b << 15; return *this; }
  FastRec& clr() { hi |= 1u << 23; return *this; }
  FastRec& inc() { hi |= 1u << 24; return *this; }
  FastRec& setaux() { hi |= 1u << 25; return *this; }
  FastRec& setjr(uint32_t v) { hi |= (1u << 26) | (v << 27); return *this; }
  FastRec& subreg() { lo |= 1u << 29; return *this; }
  FastRec& popnode() { hi |= 1u << 30; return *this; }
};

class FastTableBuilder {
 public:
  explicit FastTableBuilder(FastTables& f) : F(f) {
    std::memset(&F, 0, sizeof(F));
    std::memset(set_, 0, sizeof(set_));
    recs_.push_back({FA_BAIL, 0});  // record 0: bail
    index_[{FA_BAIL, 0}] = 0;
  }
  void on(uint32_t st, uint32_t cls, const FastRec& r) {
    if (set_[st][cls]) return;  // the first rule of a row wins, so defaults come last
    set_[st][cls] = 1;
    F.trans[st][cls] = rec(r);
  }
  void on(uint32_t st, std::initializer_list<uint32_t> classes, const FastRec& r) {
    for (uint32_t c : classes) on(st, c, r);
  }
  void bail(uint32_t st, std::initializer_list<uint32_t> classes) {
    for (uint32_t c : classes) {
      if (set_[st][c]) continue;
      set_[st][c] = 1;
      F.trans[st][c] = 0;
    }
  }
  void words(uint32_t st, const FastRec& r) {  // every word class: must_parse_identifier_string accepts any word
    for (uint32_t c = FC_FIRST_WORD; c < FC_COUNT; c++) on(st, c, r);
  }
  void ident(uint32_t st, const FastRec& r) {
    words(st, r);
    on(st, FC_DELIM, r);
  }
  void otherwise(uint32_t st, const FastRec& r) {
    for (uint32_t c = 0; c < FC_COUNT; c++) on(st, c, r);
  }
  void finish() {
    if (recs_.size() > FAST_MAX_REC) throw std::runtime_error("fast parser: too many transition records");
    for (size_t i = 0; i < recs_.size(); i++) {
      const uint32_t lo = recs_[i].first;
      F.rec_lo[i] = lo;
      F.rec_hi[i] = recs_[i].second;
      F.rec_hdr[i] = ((lo >> 15) & 255u) | (((lo >> 23) & 15u) << 8) | (((lo >> 27) & 1u) << 16);
    }
  }

 private:
  uint8_t rec(const FastRec& r) {
    const std::pair<uint32_t, uint32_t> k{r.lo, r.hi};
    auto it = index_.find(k);
    if (it != index_.end()) return it->second;
    const uint8_t i = (uint8_t)recs_.size();
    recs_.push_back(k);
    index_[k] = i;
    return i;
  }
  FastTables& F;
  uint8_t set_[FS_COUNT][FC_COUNT];
  std::vector<std::pair<uint32_t, uint32_t>> recs_;
  std::map<std::pair<uint32_t, uint32_t>, uint8_t> index_;
};

inline uint16_t fast_optok_entry(uint32_t ty) {
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
    case NUTDB_TT_LBracket: power = P_Access; bail = 1; break;  // index access: the wide pass (the narrow one declines)
    default: break;
  }
  return (uint16_t)(power | (op << 4) | (bail << 12));
}
inline uint16_t fast_opkw_entry(uint32_t kw) {
  uint32_t power = P_Terminator, op = 0, bail = 0;
  switch (kw) {
    case KW_AND: power = P_And; op = 11; break;
    case KW_OR: power = P_Or; op = 12; break;
    case KW_XOR: power = P_Xor; op = 13; break;
    case KW_IN: power = P_Comparison; op = 18; break;
    case KW_LIKE: power = P_Comparison; op = 14; break;
    case KW_ILIKE: power = P_Comparison; op = 16; break;
    // handled by the operator phase itself (FastParser::SPEC_*): NOT IN/LIKE/ILIKE/BETWEEN, IS [NOT] NULL, BETWEEN
    case KW_NOT: power = P_Not; bail = 2; break;
    case KW_IS: power = P_Comparison; bail = 3; break;
    case KW_BETWEEN: power = P_Between; bail = 4; break;
    default: break;
  }
  return (uint16_t)(power | (op << 4) | (bail << 12));
}
inline uint8_t fast_type_class(uint32_t ty) {
  switch (ty) {
    case NUTDB_TT_EOF: return FC_EOF;
    case NUTDB_TT_SemiColon: return FC_SEMI;
    case NUTDB_TT_Comma: return FC_COMMA;
    case NUTDB_TT_LParen: return FC_LPAREN;
    case NUTDB_TT_RParen: return FC_RPAREN;
    case NUTDB_TT_Mul: return FC_MUL;
    case NUTDB_TT_Plus: return FC_PLUS;
    case NUTDB_TT_Minus: return FC_MINUS;
    case NUTDB_TT_Eq: case NUTDB_TT_NotEq: case NUTDB_TT_Gt: case NUTDB_TT_Lt: case NUTDB_TT_GtEq: case NUTDB_TT_LtEq:
    case NUTDB_TT_BitOr: case NUTDB_TT_BitXor: case NUTDB_TT_BitAnd: case NUTDB_TT_BitLShift: case NUTDB_TT_BitRShift:
    case NUTDB_TT_Div: case NUTDB_TT_Mod:
      return FC_BINOP;
    case NUTDB_TT_LBracket: return FC_LBRACKET;
    case NUTDB_TT_LBrace: return FC_LBRACE;
    case NUTDB_TT_BitNot: return FC_BITNOT;
    case NUTDB_TT_IntegerLiteral: return FC_INT;
    case NUTDB_TT_HexLiteral: return FC_HEX;
    case NUTDB_TT_FloatLiteral: return FC_FLOAT;
    case NUTDB_TT_RawStringLiteral: return FC_RAWSTR;
    case NUTDB_TT_EscapedSQStringLiteral: return FC_ESQ;
    case NUTDB_TT_EscapedDQStringLiteral: return FC_EDQ;
    case NUTDB_TT_DelimitedIdentifier: return FC_DELIM;
    case NUTDB_TT_Dot: return FC_DOT;
    default: return FC_OTHER;
  }
}
inline uint8_t fast_keyword_class(uint32_t kw) {
  if (kw >= KW_INT8 && kw <= KW_NULLABLE) return FC_DTYPE;
  switch (kw) {
    case KW_TRUE: return FC_TRUE;
    case KW_FALSE: return FC_FALSE;
    case KW_NULL: return FC_NULL;
    case KW_NOT: return FC_NOT;
    case KW_IF: return FC_IF;
    case KW_INTERVAL: return FC_BADPFX;
    case KW_CASE: return FC_CASE;
    case KW_AND: case KW_OR: case KW_XOR: case KW_IN: case KW_LIKE: case KW_ILIKE: return FC_KWBINOP;
    case KW_IS: case KW_BETWEEN: return FC_ISBETWEEN;
    case KW_FROM: return FC_FROM;
    case KW_WHERE: return FC_WHERE;
    case KW_GROUP: return FC_GROUP;
    case KW_BY: return FC_BY;
    case KW_HAVING: return FC_HAVING;
    case KW_ORDER: return FC_ORDER;
    case KW_LIMIT: return FC_LIMIT;
    case KW_OFFSET: return FC_OFFSET;
    case KW_WITH: return FC_WITH;
    case KW_TIES: return FC_TIES;
    case KW_AS: return FC_AS;
    case KW_DESC: return FC_DESC;
    case KW_INTO: return FC_INTO;
    case KW_VALUES: return FC_VALUES;
    case KW_TABLE: return FC_TABLE;
    case KW_EXISTS: return FC_EXISTS;
    case KW_DEFAULT: return FC_DEFAULT;
    case KW_COMMENT: return FC_COMMENT;
    case KW_PRIMARY: return FC_PRIMARY;
    case KW_KEY: return FC_KEY;
    case KW_PARTITION: return FC_PARTITION;
    case KW_DISTINCT: return FC_DISTINCT;
    case KW_UNION: case KW_INTERSECT: case KW_EXCEPT: return FC_SETOP;
    case KW_JOIN: return FC_JOIN;
    case KW_INNER: return FC_INNER;
    case KW_FULL: return FC_FULL;
    case KW_LEFT: return FC_LEFT;
    case KW_RIGHT: return FC_RIGHT;
    case KW_OUTER: return FC_OUTER;
    case KW_SEMI: return FC_KSEMI;
    case KW_ANTI: return FC_KANTI;
    case KW_ON: return FC_ON;
    case KW_USING: return FC_USING;
    case KW_INDEX: return FC_INDEX;
    case KW_CONSTRAINT: return FC_CONSTRAINT;
    case KW_CHECK: return FC_CHECK;
    case KW_VIEW: return FC_VIEW;
    case KW_UPDATE: return FC_UPDATE;
    case KW_SELECT: return FC_SELECT;
    default: return FC_WORD;  // an identifier, or a keyword that plays no part in this subset
  }
}

inline void fast_tables_build(FastTables& F) {
  static_assert(FS_COUNT <= 128 && FC_COUNT <= 255, "state / class fields");
  FastTableBuilder B(F);
  for (uint32_t i = 0; i < 64; i++) {
    F.op[i] = fast_optok_entry(i);
    F.cls[i] = fast_type_class(i);
  }
  for (uint32_t i = 0; i < 128; i++) {
    F.op[64 + i] = fast_opkw_entry(i);
    F.cls[64 + i] = fast_keyword_class(i);
  }
  // FastTableBuilder::words() / ident() rely on the word classes being exactly [FC_FIRST_WORD, FC_COUNT)
  for (uint32_t i = 0; i < 64; i++)
    if (F.cls[i] >= FC_FIRST_WORD) throw std::runtime_error("fast parser: a token-type class lies in the word-class range");
  for (uint32_t i = 0; i < 128; i++)
    if (F.cls[64 + i] < FC_FIRST_WORD) throw std::runtime_error("fast parser: a keyword class lies below the word-class range");
  auto R = [] { return FastRec(); };
  const FastRec skip_to_operand = R().adv().to(FS_X_OPND);

  // ---- operand position: must_parse_expr_prefix (mod.rs:1222-1347) ----
  // SELECT [DISTINCT]: must_parse_query_clause_distinct (mod.rs:349-360); DISTINCT ON (..) goes to the automaton.
  // The DISTINCT node has no children (its own index is its subtree start); the select list's node starts after it.
  B.on(FS_SEL0, FC_DISTINCT, R().pre_m1().node_m1(NUTDB_NK_DISTINCT).adv().post_m0().to(FS_SEL_D));
  B.bail(FS_SEL_D, {FC_ON});
  for (uint32_t st : {(uint32_t)FS_X_OPND, (uint32_t)FS_SEL0, (uint32_t)FS_SEL_D}) {
    B.on(st, FC_INT, R().cur().check(FK_INT_W2).leaf(NUTDB_NK_LIT_INT).adv().to(FS_X_OPER));
    B.on(st, FC_HEX, R().cur().check(FK_INT_W2).leaf(NUTDB_NK_LIT_INT, 0, 1).adv().to(FS_X_OPER));
    B.on(st, FC_FLOAT, R().cur().leaf(NUTDB_NK_LIT_FLOAT).adv().to(FS_X_OPER));
    B.on(st, FC_RAWSTR, R().cur().leaf(NUTDB_NK_LIT_STR, 0).adv().to(FS_X_OPER));
    B.on(st, FC_ESQ, R().cur().check(FK_STR).leaf(NUTDB_NK_LIT_STR, 1).adv().to(FS_X_OPER));
    B.on(st, FC_EDQ, R().cur().check(FK_STR).leaf(NUTDB_NK_LIT_STR, 2).adv().to(FS_X_OPER));
    B.on(st, FC_MUL, R().cur().leaf(NUTDB_NK_IDENT, 1).adv().to(FS_X_OPER));
    B.on(st, FC_TRUE, R().cur().leaf0(NUTDB_NK_LIT_BOOL, 1).adv().to(FS_X_OPER));
    B.on(st, FC_FALSE, R().cur().leaf0(NUTDB_NK_LIT_BOOL, 0).adv().to(FS_X_OPER));
    B.on(st, FC_NULL, R().cur().leaf0(NUTDB_NK_LIT_NULL, 0).adv().to(FS_X_OPER));
    B.on(st, FC_NOT, R().act(FA_PUSH).emit(FE_NONE, 1 | (15 << 4)));  // FastParser::E_NOT with power 15
    B.on(st, FC_CASE, R().act(FA_CASE));
    // the wide pass only (these actions make the narrow pass decline): IF .. END, [array], {map}, prefix ~ push the
    // stack entry named in the record's kind field (FastParser::E_IF = 9, E_BRACKET = 10, E_MAP = 11, E_BITNOT = 2
    // with power 15); INTERVAL n unit
    B.on(st, FC_IF, R().act(FA_PUSH).emit(FE_NONE, 9));
    B.on(st, FC_LBRACKET, R().act(FA_PUSH).emit(FE_NONE, 10));
    B.on(st, FC_LBRACE, R().act(FA_PUSH).emit(FE_NONE, 11));
    B.on(st, FC_BITNOT, R().act(FA_PUSH).emit(FE_NONE, 2 | (15 << 4)));
    B.on(st, FC_BADPFX, R().act(FA_INTERVAL));
    B.ident(st, R().act(FA_IDENT));
    B.on(st, FC_MINUS, R().act(FA_NEG));
    B.on(st, FC_PLUS, R().adv().to(FS_X_OPND));  // prefix plus is dropped (mod.rs:1270)
    B.on(st, FC_LPAREN, R().act(FA_OPEN));
  }
  // FS_X_OPER has no row: the operator phase of the loop handles it and moves on to FS_AFTER + context.

  // ---- SELECT list (mod.rs:279-330), aliases (:563-578) ----
  const FastRec alias = R().leaf(NUTDB_NK_ALIAS).adv();
  const FastRec cols = R().node_m0(NUTDB_NK_COLS).to(FS_FROM);
  B.on(FS_AFTER + C_SEL_ITEM, FC_AS, R().adv().to(FS_SEL_ALIAS));
  B.ident(FS_SEL_ALIAS, FastRec(alias).to(FS_SEL_ITEM2));
  for (uint32_t st : {(uint32_t)(FS_AFTER + C_SEL_ITEM), (uint32_t)FS_SEL_ITEM2}) {
    B.on(st, FC_COMMA, skip_to_operand);
    B.otherwise(st, cols);
  }
  // ---- the clauses after the select list (mod.rs:190-203, :331-544): stage k may still see clause k.. ----
  auto clauses = [&](uint32_t st, int stage) {
    if (stage <= 1) B.on(st, FC_WHERE, R().adv().post_m1().ctx(C_WHERE).to(FS_X_OPND));
    if (stage <= 2) B.on(st, FC_GROUP, R().adv().to(FS_GROUP_BY));
    if (stage <= 3) B.on(st, FC_HAVING, R().adv().post_m1().ctx(C_HAVING).to(FS_X_OPND));
    if (stage <= 4) B.on(st, FC_ORDER, R().adv().to(FS_ORDER_BY));
    if (stage <= 5) B.on(st, FC_LIMIT, R().adv().post_m1().to(FS_LIM1));
    B.otherwise(st, R().node0(NUTDB_NK_QUERY_BODY).to(FS_END_SEL));
  };
  // [FROM name [AS a]] (must_parse_query_source, mod.rs:546-569: a plain table name here)
  B.on(FS_FROM, FC_FROM, R().adv().post_m1().to(FS_SRC));
  B.bail(FS_FROM, {FC_JOIN});
  clauses(FS_FROM, 1);
  B.bail(FS_SRC, {FC_TRUE, FC_FALSE, FC_NULL, FC_NOT, FC_IF, FC_BADPFX, FC_CASE});
  B.on(FS_SRC, FC_LPAREN, R().act(FA_SRC_SUBQ));  // FROM (select ..) [AS alias]: the wide pass
  B.words(FS_SRC, R().look(FL_NOLP).leaf(NUTDB_NK_IDENT).adv().to(FS_SRC2));  // no table function
  B.on(FS_SRC, FC_DELIM, R().leaf(NUTDB_NK_IDENT).adv().to(FS_SRC2));
  // `db.table`: the reference keeps the table and DROPS the qualifier (mod.rs:549-562): the node just emitted is withdrawn
  B.on(FS_SRC2, FC_DOT, R().adv().popnode().to(FS_SRC_Q));
  B.ident(FS_SRC_Q, R().leaf(NUTDB_NK_IDENT).adv().to(FS_SRC2B));
  // the source is an expression: anything with infix power continues it (mod.rs:1212-1216)
  for (uint32_t st : {(uint32_t)FS_SRC2, (uint32_t)FS_SRC2B}) {
    B.bail(st, {FC_BINOP, FC_MUL, FC_PLUS, FC_MINUS, FC_LBRACKET, FC_KWBINOP, FC_NOT, FC_ISBETWEEN, FC_DOT});
    B.on(st, FC_AS, R().adv().to(FS_SRC_ALIAS));
    B.otherwise(st, R().node_m1(NUTDB_NK_FROM).to(FS_CL1));
  }
  B.ident(FS_SRC_ALIAS, FastRec(alias).to(FS_SRC3));
  B.otherwise(FS_SRC3, R().node_m1(NUTDB_NK_FROM).to(FS_CL1));
  // joins (try_parse_query_clause_join, mod.rs:376-431): [INNER | FULL [OUTER] | LEFT|RIGHT [SEMI|ANTI|OUTER]] JOIN
  // source (ON expr | USING (identifiers)); the join type waits in a register for the JOIN node
  const uint32_t JOINS = FS_CL1;
  B.on(JOINS, FC_INNER, R().adv().setjr(0).to(FS_J_KW));
  B.on(JOINS, FC_FULL, R().adv().setjr(1).to(FS_J_OUTER));
  B.on(JOINS, FC_LEFT, R().adv().setjr(2).to(FS_J_LEFT));
  B.on(JOINS, FC_RIGHT, R().adv().setjr(3).to(FS_J_RIGHT));
  B.on(JOINS, FC_JOIN, R().adv().setjr(0).post_m1().to(FS_J_SRC));
  B.on(FS_J_LEFT, FC_KSEMI, R().adv().setjr(4).to(FS_J_KW));
  B.on(FS_J_LEFT, FC_KANTI, R().adv().setjr(6).to(FS_J_KW));
  B.on(FS_J_RIGHT, FC_KSEMI, R().adv().setjr(5).to(FS_J_KW));
  B.on(FS_J_RIGHT, FC_KANTI, R().adv().setjr(7).to(FS_J_KW));
  for (uint32_t st : {(uint32_t)FS_J_LEFT, (uint32_t)FS_J_RIGHT, (uint32_t)FS_J_OUTER}) B.on(st, FC_OUTER, R().adv().to(FS_J_KW));
  for (uint32_t st : {(uint32_t)FS_J_LEFT, (uint32_t)FS_J_RIGHT, (uint32_t)FS_J_OUTER, (uint32_t)FS_J_KW})
    B.on(st, FC_JOIN, R().adv().post_m1().to(FS_J_SRC));
  B.bail(FS_J_SRC, {FC_TRUE, FC_FALSE, FC_NULL, FC_NOT, FC_IF, FC_BADPFX, FC_CASE});
  B.words(FS_J_SRC, R().look(FL_NOLP).leaf(NUTDB_NK_IDENT).adv().to(FS_J_SRC2));
  B.on(FS_J_SRC, FC_DELIM, R().leaf(NUTDB_NK_IDENT).adv().to(FS_J_SRC2));
  B.on(FS_J_SRC2, FC_DOT, R().adv().popnode().to(FS_J_SRC_Q));
  B.ident(FS_J_SRC_Q, R().leaf(NUTDB_NK_IDENT).adv().to(FS_J_SRC2B));
  for (uint32_t st : {(uint32_t)FS_J_SRC2, (uint32_t)FS_J_SRC2B}) B.on(st, FC_AS, R().adv().to(FS_J_ALIAS));
  B.ident(FS_J_ALIAS, FastRec(alias).to(FS_J_ONUSING));
  for (uint32_t st : {(uint32_t)FS_J_SRC2, (uint32_t)FS_J_SRC2B, (uint32_t)FS_J_ONUSING}) {
    B.on(st, FC_ON, R().adv().ctx(C_JOIN_ON).to(FS_X_OPND));
    B.on(st, FC_USING, R().adv().to(FS_J_U_LP));
  }
  B.otherwise(FS_AFTER + C_JOIN_ON, R().node_m1(NUTDB_NK_JOIN).subreg().to(JOINS));
  B.on(FS_J_U_LP, FC_LPAREN, R().adv().to(FS_J_U_ID));
  B.ident(FS_J_U_ID, R().look(FL_NODOT).leaf(NUTDB_NK_IDENT).adv().to(FS_J_U_SEP));  // must_parse_identifier (mod.rs:1525)
  B.on(FS_J_U_ID, FC_MUL, R().leaf(NUTDB_NK_IDENT, 1).adv().to(FS_J_U_SEP));
  B.on(FS_J_U_SEP, FC_COMMA, R().adv().to(FS_J_U_ID));
  B.on(FS_J_U_SEP, FC_RPAREN, R().node_m1(NUTDB_NK_JOIN, 0, 1).subreg().adv().to(JOINS));
  clauses(FS_CL1, 1);
  clauses(FS_CL2, 2);
  clauses(FS_CL3, 3);
  clauses(FS_CL4, 4);
  clauses(FS_CL5, 5);
  B.on(FS_GROUP_BY, FC_BY, R().adv().post_m1().ctx(C_GROUP_ITEM).to(FS_X_OPND));
  B.on(FS_ORDER_BY, FC_BY, R().adv().post_m1().ctx(C_ORDER_ITEM).to(FS_X_OPND));
  B.otherwise(FS_AFTER + C_WHERE, R().node_m1(NUTDB_NK_WHERE).to(FS_CL2));
  B.otherwise(FS_AFTER + C_HAVING, R().node_m1(NUTDB_NK_HAVING).to(FS_CL4));
  // GROUP BY items
  B.on(FS_AFTER + C_GROUP_ITEM, FC_AS, R().adv().to(FS_GRP_ALIAS));
  B.ident(FS_GRP_ALIAS, FastRec(alias).to(FS_GRP_ITEM2));
  for (uint32_t st : {(uint32_t)(FS_AFTER + C_GROUP_ITEM), (uint32_t)FS_GRP_ITEM2}) {
    B.on(st, FC_COMMA, skip_to_operand);
    B.otherwise(st, R().node_m1(NUTDB_NK_GROUPBY).to(FS_CL3));
  }
  // ORDER BY items: DESC only, the reference never accepts ASC (mod.rs:491-496)
  B.on(FS_AFTER + C_ORDER_ITEM, FC_AS, R().adv().to(FS_ORD_ALIAS));
  B.ident(FS_ORD_ALIAS, FastRec(alias).to(FS_ORD_ITEM2));
  for (uint32_t st : {(uint32_t)(FS_AFTER + C_ORDER_ITEM), (uint32_t)FS_ORD_ITEM2, (uint32_t)FS_ORD_ITEM3}) {
    if (st != FS_ORD_ITEM3) B.on(st, FC_DESC, R().leaf0(NUTDB_NK_ORDER_DESC).adv().to(FS_ORD_ITEM3));
    B.on(st, FC_COMMA, skip_to_operand);
    B.otherwise(st, R().node_m1(NUTDB_NK_ORDERBY).to(FS_CL5));
  }
  // LIMIT n [, m | OFFSET m] [WITH TIES] (mod.rs:503-544)
  auto number = [&](uint32_t st, uint32_t next) {  // must_parse_integer_literal (mod.rs:1815) -> NK_NUM
    B.on(st, FC_INT, R().check(FK_INT_W1).leaf(NUTDB_NK_NUM, 0, 0).adv().to(next));
    B.on(st, FC_HEX, R().check(FK_INT_W1).leaf(NUTDB_NK_NUM, 0, 1).adv().to(next));
  };
  number(FS_LIM1, FS_LIM2);
  B.on(FS_LIM2, FC_COMMA, R().adv().to(FS_LIM3A));
  B.on(FS_LIM2, FC_OFFSET, R().adv().to(FS_LIM3B));
  number(FS_LIM3A, FS_LIM4A);
  number(FS_LIM3B, FS_LIM4B);
  const uint32_t lim_end[3] = {FS_LIM2, FS_LIM4A, FS_LIM4B}, ties[3] = {FS_TIES0, FS_TIES1, FS_TIES2};
  for (uint32_t k = 0; k < 3; k++) {
    B.on(lim_end[k], FC_WITH, R().adv().to(ties[k]));
    B.otherwise(lim_end[k], R().node_m1(NUTDB_NK_LIMIT, k, 0).to(FS_BODY));
    B.on(ties[k], FC_TIES, R().node_m1(NUTDB_NK_LIMIT, k, 1).adv().to(FS_BODY));
  }
  B.otherwise(FS_BODY, R().node0(NUTDB_NK_QUERY_BODY).to(FS_END_SEL));
  B.on(FS_END_SEL, FC_SETOP, R().act(FA_SETOP));  // set operations (mod.rs:250-267): the wide pass
  B.on(FS_END_SEL, FC_RPAREN, R().act(FA_SUBQ_END));  // the `)` of a parenthesised subquery (wide pass)
  B.otherwise(FS_END_SEL, R().node0(NUTDB_NK_STMT_SELECT).to(FS_FINAL));
  // WITH name AS (select ..) {, name AS (select ..)} SELECT .. (must_parse_query_clause_with, mod.rs:327-347): wide pass.
  // The WITH node covers the common table expressions (m1 = where they begin); the select list starts behind it.
  B.ident(FS_WITH0, R().leaf(NUTDB_NK_NAME).adv().to(FS_WITH_AS));
  B.on(FS_WITH_AS, FC_AS, R().adv().to(FS_WITH_LP));
  B.on(FS_WITH_LP, FC_LPAREN, R().act(FA_SRC_SUBQ).emit(FE_NONE, 1));
  B.on(FS_WITH_SEP, FC_COMMA, R().adv().to(FS_WITH0));
  B.on(FS_WITH_SEP, FC_SELECT, R().node_m1(NUTDB_NK_WITH).adv().post_m0().to(FS_SEL0));
  // statement-final position (mod.rs:165-172)
  B.on(FS_FINAL, {FC_EOF, FC_SEMI}, R().act(FA_ACCEPT));

  // ---- INSERT INTO name [(names)] VALUES rows (mod.rs:589-670) ----
  const FastRec name = R().leaf(NUTDB_NK_NAME).adv();
  B.on(FS_INS0, FC_INTO, R().adv().to(FS_INS_NAME));
  B.ident(FS_INS_NAME, FastRec(name).to(FS_INS_AFTER_NAME));
  B.on(FS_INS_AFTER_NAME, FC_LPAREN, R().adv().to(FS_INS_COL));
  B.ident(FS_INS_COL, FastRec(name).to(FS_INS_COL_SEP));
  B.on(FS_INS_COL_SEP, FC_COMMA, R().adv().to(FS_INS_COL));
  B.on(FS_INS_COL_SEP, FC_RPAREN, R().adv().to(FS_INS_VALUES));
  for (uint32_t st : {(uint32_t)FS_INS_AFTER_NAME, (uint32_t)FS_INS_VALUES})
    B.on(st, FC_VALUES, R().adv().post_m0().to(FS_INS_ROW0));
  B.on(FS_INS_ROW0, FC_LPAREN, R().adv().post_m1().ctx(C_INS_VALUE).to(FS_X_OPND));
  B.on(FS_AFTER + C_INS_VALUE, FC_COMMA, R().adv().inc().to(FS_X_OPND));
  B.on(FS_AFTER + C_INS_VALUE, FC_RPAREN, R().act(FA_ROWEND));
  B.on(FS_INS_AFTER_ROW, FC_COMMA, R().adv().to(FS_INS_ROWN));
  B.otherwise(FS_INS_AFTER_ROW, R().node_m0(NUTDB_NK_ROWS).to(FS_INS_END));
  B.on(FS_INS_ROWN, FC_LPAREN, R().adv().post_m1().to(FS_X_OPND));
  B.otherwise(FS_INS_END, R().node0(NUTDB_NK_STMT_INSERT).to(FS_FINAL));

  // ---- CREATE VIEW [IF NOT EXISTS] name {UPDATE BY strategy | PRIMARY KEY es | ORDER BY es | PARTITION BY e | COMMENT s}
  //      AS query (must_parse_view_definition, mod.rs:807-911): wide pass (the AS action is its own).  seen bit 1 =
  //      UPDATE BY (AS needs it), the other bits as for tables ----
  B.on(FS_CRE0, FC_VIEW, R().adv().to(FS_CRV1));
  B.on(FS_CRV1, FC_IF, R().adv().to(FS_CRV_IF1));
  B.on(FS_CRV_IF1, FC_NOT, R().adv().to(FS_CRV_IF2));
  B.on(FS_CRV_IF2, FC_EXISTS, R().adv().setaux().to(FS_CRV_NAME));
  for (uint32_t st : {(uint32_t)FS_CRV1, (uint32_t)FS_CRV_NAME}) B.ident(st, R().leaf(NUTDB_NK_NAME).adv().to(FS_VIEW_ATTRS));
  B.on(FS_VIEW_ATTRS, FC_UPDATE, R().bit(1).adv().to(FS_V_UPD));
  B.on(FS_VIEW_ATTRS, FC_PRIMARY, R().bit(16).adv().to(FS_V_KEY));
  B.on(FS_VIEW_ATTRS, FC_ORDER, R().bit(32).adv().to(FS_V_ORDER_BY));
  B.on(FS_VIEW_ATTRS, FC_PARTITION, R().bit(64).adv().to(FS_V_PART_BY));
  B.on(FS_VIEW_ATTRS, FC_COMMENT, R().bit(128).adv().to(FS_V_COMMENT));
  B.on(FS_VIEW_ATTRS, FC_AS, R().act(FA_SRC_SUBQ).emit(FE_NONE, 2));
  B.on(FS_V_UPD, FC_BY, R().adv().to(FS_V_STRAT));
  B.ident(FS_V_STRAT, R().leaf(NUTDB_NK_STRATEGY).adv().to(FS_VIEW_ATTRS));
  B.on(FS_V_KEY, FC_KEY, R().adv().post_m1().ctx(C_V_PK_ITEM).to(FS_X_OPND));
  B.on(FS_V_ORDER_BY, FC_BY, R().adv().post_m1().ctx(C_V_ORDER_ITEM).to(FS_X_OPND));
  B.on(FS_V_PART_BY, FC_BY, R().adv().post_m1().ctx(C_V_PART).to(FS_X_OPND));
  B.on(FS_AFTER + C_V_PK_ITEM, FC_COMMA, skip_to_operand);
  B.otherwise(FS_AFTER + C_V_PK_ITEM, R().node_m1(NUTDB_NK_ATTR_PK).to(FS_VIEW_ATTRS));
  B.on(FS_AFTER + C_V_ORDER_ITEM, FC_COMMA, skip_to_operand);
  B.otherwise(FS_AFTER + C_V_ORDER_ITEM, R().node_m1(NUTDB_NK_ATTR_ORDER).to(FS_VIEW_ATTRS));
  B.otherwise(FS_AFTER + C_V_PART, R().node_m1(NUTDB_NK_ATTR_PART).to(FS_VIEW_ATTRS));

  // ---- CREATE TABLE (mod.rs:689-805, :936-972) ----
  B.on(FS_CRE0, FC_TABLE, R().adv().to(FS_CRE1));
  B.on(FS_CRE1, FC_IF, R().adv().to(FS_CRE_IF1));
  B.on(FS_CRE_IF1, FC_NOT, R().adv().to(FS_CRE_IF2));
  B.on(FS_CRE_IF2, FC_EXISTS, R().adv().setaux().to(FS_CRE_NAME));
  for (uint32_t st : {(uint32_t)FS_CRE1, (uint32_t)FS_CRE_NAME}) B.ident(st, FastRec(name).to(FS_CRE_LP));
  B.on(FS_CRE_LP, FC_LPAREN, R().adv().to(FS_COL_BEGIN));
  // INDEX name indexer(..) / CONSTRAINT name CHECK expr among the columns (mod.rs:913-934)
  B.on(FS_COL_BEGIN, FC_INDEX, R().adv().pre_m0().to(FS_IDX_NAME));
  B.on(FS_COL_BEGIN, FC_CONSTRAINT, R().adv().pre_m0().to(FS_CON_NAME));
  B.ident(FS_IDX_NAME, FastRec(name).ctx(C_IDX_EXPR).to(FS_X_OPND));
  B.ident(FS_CON_NAME, FastRec(name).to(FS_CON_CHECK));
  B.on(FS_CON_CHECK, FC_CHECK, R().adv().ctx(C_CON_EXPR).to(FS_X_OPND));
  B.otherwise(FS_AFTER + C_IDX_EXPR, R().check(FK_FNCALL).node_m0(NUTDB_NK_INDEXDEF).to(FS_COL_SEP));
  B.otherwise(FS_AFTER + C_CON_EXPR, R().node_m0(NUTDB_NK_CONSTRDEF).to(FS_COL_SEP));
  B.ident(FS_COL_BEGIN, FastRec(name).pre_m0().clr().to(FS_DT));
  B.on(FS_DT, FC_DTYPE, R().act(FA_DTYPE));
  B.otherwise(FS_DT_END, R().act(FA_DTEND));
  // column attributes, each at most once (mod.rs:936-972)
  B.on(FS_COL_ATTRS, FC_DEFAULT, R().bit(1).adv().post_m1().ctx(C_COL_DEFAULT).to(FS_X_OPND));
  B.on(FS_COL_ATTRS, FC_COMMENT, R().bit(2).adv().to(FS_COL_COMMENT));
  for (uint32_t c = FC_FIRST_WORD; c < FC_COUNT; c++) B.bail(FS_COL_ATTRS, {c});  // any other word: the automaton
  B.otherwise(FS_COL_ATTRS, R().node_m0(NUTDB_NK_COLDEF).to(FS_COL_SEP));
  auto string_lit = [&](uint32_t st, uint32_t next) {  // must_parse_string_literal (mod.rs:1833) -> NK_STR
    B.on(st, FC_RAWSTR, R().leaf(NUTDB_NK_STR, 0).adv().to(next));
    B.on(st, FC_ESQ, R().check(FK_STR).leaf(NUTDB_NK_STR, 1).adv().to(next));
    B.on(st, FC_EDQ, R().check(FK_STR).leaf(NUTDB_NK_STR, 2).adv().to(next));
  };
  string_lit(FS_COL_COMMENT, FS_COL_ATTRS);
  string_lit(FS_V_COMMENT, FS_VIEW_ATTRS);
  B.otherwise(FS_AFTER + C_COL_DEFAULT, R().node_m1(NUTDB_NK_ATTR_DEFAULT).to(FS_COL_ATTRS));
  B.on(FS_COL_SEP, FC_COMMA, R().adv().to(FS_COL_BEGIN));
  B.on(FS_COL_SEP, FC_RPAREN, R().adv().to(FS_TBL_ATTRS));
  // table attributes after the column list, each at most once (mod.rs:746-803)
  B.on(FS_TBL_ATTRS, FC_PRIMARY, R().bit(16).adv().to(FS_TBL_KEY));
  B.on(FS_TBL_ATTRS, FC_ORDER, R().bit(32).adv().to(FS_TBL_ORDER_BY));
  B.on(FS_TBL_ATTRS, FC_PARTITION, R().bit(64).adv().to(FS_TBL_PART_BY));
  B.on(FS_TBL_ATTRS, FC_COMMENT, R().bit(128).adv().to(FS_TBL_COMMENT));
  for (uint32_t c = FC_FIRST_WORD; c < FC_COUNT; c++) B.bail(FS_TBL_ATTRS, {c});
  B.otherwise(FS_TBL_ATTRS, R().node0(NUTDB_NK_TABLEDEF).to(FS_CRE_END));
  B.on(FS_TBL_KEY, FC_KEY, R().adv().post_m1().ctx(C_TBL_PK_ITEM).to(FS_X_OPND));
  B.on(FS_TBL_ORDER_BY, FC_BY, R().adv().post_m1().ctx(C_TBL_ORDER_ITEM).to(FS_X_OPND));
  B.on(FS_TBL_PART_BY, FC_BY, R().adv().post_m1().ctx(C_TBL_PART).to(FS_X_OPND));
  string_lit(FS_TBL_COMMENT, FS_TBL_ATTRS);
  B.on(FS_AFTER + C_TBL_PK_ITEM, FC_COMMA, skip_to_operand);
  B.otherwise(FS_AFTER + C_TBL_PK_ITEM, R().node_m1(NUTDB_NK_ATTR_PK).to(FS_TBL_ATTRS));
  B.on(FS_AFTER + C_TBL_ORDER_ITEM, FC_COMMA, skip_to_operand);
  B.otherwise(FS_AFTER + C_TBL_ORDER_ITEM, R().node_m1(NUTDB_NK_ATTR_ORDER).to(FS_TBL_ATTRS));
  B.otherwise(FS_AFTER + C_TBL_PART, R().node_m1(NUTDB_NK_ATTR_PART).to(FS_TBL_ATTRS));
  B.otherwise(FS_CRE_END, R().node0(NUTDB_NK_STMT_CREATE).auxreg().to(FS_FINAL));
  B.finish();
}

}  // namespace npar
