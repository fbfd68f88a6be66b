// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_lex.hpp header).
//
// Rust `{:?}` (derive(Debug)) rendering of the oracle AST, `Display` rendering of ParseError
// (src/parser/error.rs:8-57, tokenizer/error.rs:14-30, tokenizer/token.rs Display via derive_more),
// and serialisation of the tree into the flat node format of include/nutdb_gpu.h.
#pragma once
#include "oracle_parse.hpp"

namespace ora {

// TokenType Display: variant name, with two overrides (tokenizer/token.rs:20,23)
inline const char* token_type_name(TT t) {
  static const char* const names[] = {
      "KeywordOrIdentifier", "DelimitedIdentifier", "ConfigIdentifier", "QueryParameter", "RawStringLiteral",
      "EscapedSingleQuotedStringLiteral", "EscapedDoubleQuotedStringLiteral", "IntegerLiteral", "FloatLiteral",
      "HexLiteral", "Comma", "Dot", "Colon", "SemiColon", "Plus", "Minus", "Mul", "Div", "Mod", "Eq", "NotEq", "Lt",
      "Gt", "LtEq", "GtEq", "LParen", "RParen", "LBracket", "RBracket", "LBrace", "RBrace", "BitAnd", "BitOr",
      "BitXor", "BitNot", "BitLShift", "BitRShift", "Comment", "Whitespace", "EOF", "POISON"};
  return t <= NUTDB_TT_POISON ? names[t] : "?";
}

inline std::string u128_to_string(u128 v) {
  if (v == 0) return "0";
  std::string s;
  while (v) {
    s.push_back((char)('0' + (int)(v % 10)));
    v /= 10;
  }
  return std::string(s.rbegin(), s.rend());
}

// <str as Debug>: escape_debug without escaping single quotes.  Non-ASCII is passed through
// (Rust additionally escapes unprintable / grapheme-extend code points; not restated).
inline void dbg_str(std::string& o, sv s) {
  o.push_back('"');
  for (unsigned char c : s) {
    switch (c) {
      case '\0': o += "\\0"; break;
      case '\t': o += "\\t"; break;
      case '\r': o += "\\r"; break;
      case '\n': o += "\\n"; break;
      case '\\': o += "\\\\"; break;
      case '"': o += "\\\""; break;
      default:
        if (c < 0x20 || c == 0x7f) {
          static const char* hx = "0123456789abcdef";
          o += "\\u{";
          if (c >= 16) o.push_back(hx[c >> 4]);
          o.push_back(hx[c & 15]);
          o += "}";
        } else {
          o.push_back((char)c);
        }
    }
  }
  o.push_back('"');
}

static const char* const UNARY_NAMES[] = {"BitwiseNot", "Not", "IsNull", "IsNotNull"};
static const char* const BINARY_NAMES[] = {
    "Plus", "Minus", "Multi", "Div", "Mod", "Gt", "Lt", "GtEq", "LtEq", "Eq", "NotEq", "And", "Or", "Xor", "Like",
    "NotLike", "ILike", "NotILike", "In", "NotIn", "IndexAccess", "BitwiseOr", "BitwiseAnd", "BitwiseXor",
    "BitwiseLeftShift", "BitwiseRightShift"};
static const char* const FN_NAMES[] = {"If", "MultiIf", "CaseWhen", "Between", "NotBetween", "Exists", "NotExists"};
static const char* const COLL_NAMES[] = {"Tuple", "Map", "Array"};
static const char* const UNIT_NAMES[] = {"Second", "Minute", "Hour", "Day", "Month", "Year"};
static const char* const JOIN_NAMES[] = {"Inner", "FullOuter", "LeftOuter", "RightOuter", "LeftSemi", "RightSemi",
                                         "LeftAnti", "RightAnti", "AsOf"};
static const char* const UNION_NAMES[] = {"UnionAll", "UnionDistinct", "Intersect", "Except"};
static const char* const SCALAR_NAMES[] = {
    "Int8", "Int16", "Int32", "Int64", "Int128", "UInt8", "UInt16", "UInt32", "UInt64", "UInt128", "Serial32",
    "Serial64", "Serial128", "USerial32", "USerial64", "USerial128", "Decimal32", "Decimal64", "Float32", "Float64",
    "Boolean", "Chars", "String", "Uuid", "Date", "Datetime"};
static const char* const COMPOUND_NAMES[] = {"Array", "Enum", "Tuple", "Map", "Dictionary", "Nullable"};

struct DebugFmt {
  std::string o;

  void query(const Query& q);
  void expr(const Expr& e);

  void opt_str(const std::optional<sv>& s) {
    if (s) {
      o += "Some(";
      dbg_str(o, *s);
      o += ")";
    } else {
      o += "None";
    }
  }
  void opt_strlit(const std::optional<StringLit>& s) {
    if (s) {
      o += "Some(";
      dbg_str(o, s->value);
      o += ")";
    } else {
      o += "None";
    }
  }
  void fncall(const Expr& e) {
    o += "FnCall { callee: ";
    if (e.sub == F_Others) {
      o += "Others(";
      dbg_str(o, e.s1);
      o += ")";
    } else {
      o += FN_NAMES[e.sub];
    }
    o += ", arguments: ";
    expr_list(e.kids);
    o += " }";
  }
  void expr_list(const std::vector<Expr>& v) {
    o += "[";
    for (size_t i = 0; i < v.size(); i++) {
      if (i) o += ", ";
      expr(v[i]);
    }
    o += "]";
  }
  void opt_expr(const std::optional<Expr>& e) {
    if (e) {
      o += "Some(";
      expr(*e);
      o += ")";
    } else {
      o += "None";
    }
  }
  void opt_expr_list(const std::optional<std::vector<Expr>>& v) {
    if (v) {
      o += "Some(";
      expr_list(*v);
      o += ")";
    } else {
      o += "None";
    }
  }
  void identifier(const Expr& e) {
    o += "Identifier { name: ";
    if (e.sub == 1) {
      o += "Wildcard";
    } else {
      o += "Word(";
      dbg_str(o, e.s1);
      o += ")";
    }
    o += ", qualifier: ";
    if (e.has_qual) {
      o += "Some(";
      dbg_str(o, e.s2);
      o += ")";
    } else {
      o += "None";
    }
    o += " }";
  }
  void query_expr(const QueryExpr& q) {
    o += "QueryExpr { inner: ";
    expr(q.inner);
    o += ", alias: ";
    opt_str(q.alias);
    o += " }";
  }
  void query_expr_list(const std::vector<QueryExpr>& v) {
    o += "[";
    for (size_t i = 0; i < v.size(); i++) {
      if (i) o += ", ";
      query_expr(v[i]);
    }
    o += "]";
  }
  void source(const QuerySource& s) {
    o += "QuerySource { inner: ";
    if (s.kind == QuerySource::TableFn) {
      o += "TableFn(";
      fncall(s.fn);
      o += ")";
    } else if (s.kind == QuerySource::Table) {
      o += "Table(";
      dbg_str(o, s.table);
      o += ")";
    } else {
      o += "Subquery(";
      query(*s.q);
      o += ")";
    }
    o += ", alias: ";
    opt_str(s.alias);
    o += " }";
  }
  void body(const QueryBody& b) {
    o += "QueryBody { with: ";
    if (b.with) {
      o += "Some(WithClause { cte_list: [";
      for (size_t i = 0; i < b.with->size(); i++) {
        if (i) o += ", ";
        o += "QueryCTE { subquery: ";
        query(*(*b.with)[i].subquery);
        o += ", alias: ";
        dbg_str(o, (*b.with)[i].alias);
        o += " }";
      }
      o += "] })";
    } else {
      o += "None";
    }
    o += ", distinct: ";
    if (b.has_distinct) {
      o += "Some(DistinctClause { columns: ";
      if (b.distinct_on) {
        o += "Some(";
        query_expr_list(*b.distinct_on);
        o += ")";
      } else {
        o += "None";
      }
      o += " })";
    } else {
      o += "None";
    }
    o += ", columns: ";
    query_expr_list(b.columns);
    o += ", from: ";
    if (b.from) {
      o += "Some(FromClause { source: ";
      source(*b.from);
      o += " })";
    } else {
      o += "None";
    }
    o += ", joins: [";
    for (size_t i = 0; i < b.joins.size(); i++) {
      const JoinClause& j = b.joins[i];
      if (i) o += ", ";
      o += "JoinClause { typ: ";
      o += JOIN_NAMES[j.typ];
      o += ", source: ";
      source(j.source);
      o += ", condition: ";
      if (j.is_using) {
        o += "Using([";
        for (size_t k = 0; k < j.using_cols.size(); k++) {
          if (k) o += ", ";
          identifier(j.using_cols[k]);
        }
        o += "])";
      } else {
        o += "On(";
        expr(*j.on);
        o += ")";
      }
      o += " }";
    }
    o += "], where: ";  // field r#where prints as `where`
    if (b.where) {
      o += "Some(WhereClause { condition: ";
      expr(*b.where);
      o += " })";
    } else {
      o += "None";
    }
    o += ", group_by: ";
    if (b.group_by) {
      o += "Some(GroupByClause { keys: ";
      query_expr_list(*b.group_by);
      o += " })";
    } else {
      o += "None";
    }
    o += ", having: ";
    if (b.having) {
      o += "Some(HavingClause { condition: ";
      expr(*b.having);
      o += " })";
    } else {
      o += "None";
    }
    o += ", order_by: ";
    if (b.order_by) {
      o += "Some(OrderByClause { keys: [";
      for (size_t i = 0; i < b.order_by->size(); i++) {
        if (i) o += ", ";
        o += "QueryOrderKey { expr: ";
        query_expr((*b.order_by)[i].expr);
        o += ", direction: ";
        o += (*b.order_by)[i].desc ? "DESC" : "ASC";
        o += " }";
      }
      o += "] })";
    } else {
      o += "None";
    }
    o += ", limit: ";
    if (b.limit) {
      o += "Some(LimitClause { size: " + std::to_string(b.limit->size) + ", offset: " +
           std::to_string(b.limit->offset) + ", with_ties: " + (b.limit->with_ties ? "true" : "false") + " })";
    } else {
      o += "None";
    }
    o += " }";
  }
  void datatype(const DataType& d) {
    if (!d.compound) {
      o += "Scalar(";
      o += SCALAR_NAMES[d.id];
      if (d.id == 16 || d.id == 17) o += " { scale: " + std::to_string(d.param) + " }";
      else if (d.id == 21) o += " { length: " + std::to_string(d.param) + " }";
      else if (d.id == 22) o += " { max_length: " + std::to_string(d.param) + " }";
      o += ")";
      return;
    }
    o += "Compound(";
    o += COMPOUND_NAMES[d.id];
    o += "(";
    if (d.id == 1) {
      o += "[";
      for (size_t i = 0; i < d.binds.size(); i++) {
        if (i) o += ", ";
        o += "EnumBind { id: " + std::to_string(d.binds[i].id) + ", literal: ";
        dbg_str(o, d.binds[i].literal.value);
        o += " }";
      }
      o += "]";
    } else if (d.id == 2) {
      o += "[";
      for (size_t i = 0; i < d.inner.size(); i++) {
        if (i) o += ", ";
        datatype(d.inner[i]);
      }
      o += "]";
    } else {
      for (size_t i = 0; i < d.inner.size(); i++) {
        if (i) o += ", ";
        datatype(d.inner[i]);
      }
    }
    o += "))";
  }
  void coldef(const ColumnDefinition& c) {
    o += "ColumnDefinition { name: ";
    dbg_str(o, c.name);
    o += ", typ: ";
    datatype(c.typ);
    o += ", default: ";
    opt_expr(c.default_);
    o += ", comment: ";
    opt_strlit(c.comment);
    o += " }";
  }
  void condef(const ConstraintDefinition& c) {
    o += "ConstraintDefinition { name: ";
    dbg_str(o, c.name);
    o += ", check: ";
    expr(c.check);
    o += " }";
  }
  void idxdef(const IndexDefinition& d) {
    o += "IndexDefinition { name: ";
    dbg_str(o, d.name);
    o += ", indexer: ";
    fncall(d.indexer);
    o += " }";
  }
  void statement(const Statement& s) {
    switch (s.k) {
      case Statement::Select:
        o += "Select(SelectStmt { query: ";
        query(*s.query);
        o += " })";
        break;
      case Statement::Explain:
        o += "Explain(ExplainStmt { query: ";
        query(*s.query);
        o += " })";
        break;
      case Statement::Insert:
        o += "Insert(InsertStmt { table_name: ";
        dbg_str(o, s.table_name);
        o += ", column_list: ";
        if (s.column_list) {
          o += "Some([";
          for (size_t i = 0; i < s.column_list->size(); i++) {
            if (i) o += ", ";
            dbg_str(o, (*s.column_list)[i]);
          }
          o += "])";
        } else {
          o += "None";
        }
        o += ", data: ";
        if (s.insert_kind == 0) {
          o += "Rows { column_size: " + std::to_string(s.column_size) + ", data: ";
          expr_list(s.rows_data);
          o += " }";
        } else if (s.insert_kind == 1) {
          o += "Subquery(";
          query(*s.query);
          o += ")";
        } else {
          o += "FnCall(";
          fncall(s.insert_fn);
          o += ")";
        }
        o += " })";
        break;
      case Statement::Alter: {
        o += "Alter(AlterStmt { alter: Alter { action: ";
        if (s.alter_action == 0) {
          o += "Add { entity: ";
          if (s.entity_kind == 0) { o += "Column("; coldef(s.col); o += ")"; }
          else if (s.entity_kind == 1) { o += "Index("; idxdef(s.idx); o += ")"; }
          else { o += "Constraint("; condef(s.con); o += ")"; }
          o += std::string(", if_not_exists: ") + (s.flag ? "true" : "false") + ", position: ";
          if (s.position == 0) o += "First";
          else if (s.position == 1) { o += "After("; dbg_str(o, s.after_name); o += ")"; }
          else o += "Last";
          o += " }";
        } else if (s.alter_action == 1) {
          o += "Drop { entity: ";
          static const char* const n[] = {"Column", "Index", "Constraint", "Partition"};
          o += n[s.entity_kind];
          o += "(";
          if (s.entity_kind == 3) dbg_str(o, s.partition.value);
          else dbg_str(o, s.entity_name);
          o += std::string("), if_exists: ") + (s.flag ? "true" : "false") + " }";
        } else {
          o += "Rename { entity: ";
          static const char* const n[] = {"Column", "Index", "Constraint", "Table"};
          o += n[s.entity_kind];
          if (s.entity_kind != 3) { o += "("; dbg_str(o, s.entity_name); o += ")"; }
          o += ", new_name: ";
          dbg_str(o, s.new_name);
          o += " }";
        }
        o += ", table_name: ";
        dbg_str(o, s.table_name);
        o += " } })";
        break;
      }
      case Statement::Create:
        o += std::string("Create(CreateStmt { if_not_exists: ") + (s.flag ? "true" : "false") + ", entity_def: ";
        if (!s.is_view) {
          const TableDefinition& t = s.table;
          o += "Table(TableDefinition { name: ";
          dbg_str(o, t.name);
          o += ", columns: [";
          for (size_t i = 0; i < t.columns.size(); i++) { if (i) o += ", "; coldef(t.columns[i]); }
          o += "], constraints: [";
          for (size_t i = 0; i < t.constraints.size(); i++) { if (i) o += ", "; condef(t.constraints[i]); }
          o += "], indexes: [";
          for (size_t i = 0; i < t.indexes.size(); i++) { if (i) o += ", "; idxdef(t.indexes[i]); }
          o += "], primary_key: ";
          opt_expr_list(t.primary_key);
          o += ", order_by: ";
          opt_expr_list(t.order_by);
          o += ", partition_by: ";
          opt_expr(t.partition_by);
          o += ", comment: ";
          opt_strlit(t.comment);
          o += " })";
        } else {
          const ViewDefinition& v = s.view;
          o += "View(ViewDefinition { name: ";
          dbg_str(o, v.name);
          o += ", strategy: ";
          dbg_str(o, v.strategy);
          o += ", primary_key: ";
          opt_expr_list(v.primary_key);
          o += ", order_by: ";
          opt_expr_list(v.order_by);
          o += ", partition_by: ";
          opt_expr(v.partition_by);
          o += ", query: ";
          query(*v.query);
          o += ", comment: ";
          opt_strlit(v.comment);
          o += " })";
        }
        o += " })";
        break;
      case Statement::Describe:
        o += "Describe(DescribeStmt { entity: ";
        if (s.entity == 2) o += "Database";
        else { o += s.entity == 0 ? "Table(" : "View("; dbg_str(o, s.name); o += ")"; }
        o += " })";
        break;
      case Statement::Drop:
      case Statement::Truncate:
        o += s.k == Statement::Drop ? "Drop(DropStmt { typ: " : "Truncate(TruncateStmt { typ: ";
        o += s.entity == 0 ? "Table" : "View";
        o += std::string(", if_exists: ") + (s.flag ? "true" : "false") + ", name: ";
        dbg_str(o, s.name);
        o += " })";
        break;
      case Statement::Optimize:
        o += "Optimize(OptimizeStmt { table_name: ";
        dbg_str(o, s.table_name);
        o += ", partition_key: ";
        opt_expr(s.partition_key);
        o += " })";
        break;
      case Statement::Set:
        o += "Set(SetStmt { config_name: ";
        dbg_str(o, s.config_name);
        o += ", value: ";
        expr(s.value);
        o += " })";
        break;
    }
  }
};

inline void DebugFmt::query(const Query& q) {
  if (!q.is_union) {
    o += "Single(";
    body(*q.body);
    o += ")";
  } else {
    o += "Union { typ: ";
    o += UNION_NAMES[q.typ];
    o += ", left: ";
    query(*q.left);
    o += ", right: ";
    query(*q.right);
    o += " }";
  }
}

inline void DebugFmt::expr(const Expr& e) {
  switch (e.k) {
    case Expr::Identifier:
      o += "Identifier(";
      identifier(e);
      o += ")";
      break;
    case Expr::QueryParameter:
      o += "QueryParameter(QueryParameter { index: " + u128_to_string(e.ival) + " })";
      break;
    case Expr::Literal:
      o += "Literal(";
      switch (e.sub) {
        case Expr::LInteger: o += "Integer(" + u128_to_string(e.ival) + ", " + (e.flag ? "true" : "false") + ")"; break;
        case Expr::LFloat: o += "Float(BigDecimal(\"" + e.dec->display() + "\"))"; break;
        case Expr::LString: o += "String("; dbg_str(o, e.str_value()); o += ")"; break;
        case Expr::LBoolean: o += std::string("Boolean(") + (e.flag ? "true" : "false") + ")"; break;
        case Expr::LInterval: o += "Interval(" + u128_to_string(e.ival) + ", " + UNIT_NAMES[e.unit] + ")"; break;
        default: o += "Null";
      }
      o += ")";
      break;
    case Expr::Collection:
      o += std::string("Collection(Collection { typ: ") + COLL_NAMES[e.sub] + ", items: ";
      expr_list(e.kids);
      o += " })";
      break;
    case Expr::UnaryOp:
      o += std::string("UnaryOp(UnaryOp { op: ") + UNARY_NAMES[e.sub] + ", operand: ";
      expr(e.kids[0]);
      o += " })";
      break;
    case Expr::BinaryOp:
      o += std::string("BinaryOp(BinaryOp { op: ") + BINARY_NAMES[e.sub] + ", left: ";
      expr(e.kids[0]);
      o += ", right: ";
      expr(e.kids[1]);
      o += " })";
      break;
    case Expr::FnCall:
      o += "FnCall(";
      fncall(e);
      o += ")";
      break;
    case Expr::Subquery:
      o += "Subquery(";
      query(*e.q);
      o += ")";
      break;
  }
}

// ------------------------------------------------------------------------------------------
// Display of errors
// ------------------------------------------------------------------------------------------
inline std::string pos_str(const Position& p) {  // utf8_iter.rs:34-38
  return "line " + std::to_string(p.line) + " col " + std::to_string(p.col);
}
inline std::string error_display(const ParseErr& e) {
  if (e.is_lex)  // error.rs:10 + tokenizer/error.rs:24-25
    return std::string("Lex Error: ") + tok_err_type_str(e.lex.t) + ": " + e.lex.ctx + " near " + pos_str(e.lex.pos);
  const SyntaxError& s = e.syn;
  std::string r = "Syntax Error: ";
  auto join = [](const std::vector<std::string>& v) {
    std::string j;
    for (size_t i = 0; i < v.size(); i++) {
      if (i) j += ", ";
      j += v[i];
    }
    return j;
  };
  switch (s.variant) {
    case NUTDB_SE_NotExpectedTokenTypes: {
      std::vector<std::string> names;
      for (TT t : s.expected_types) names.emplace_back(token_type_name(t));
      r += "expected token (" + join(names) + ") but found token " + token_type_name(s.actual_type) + " at " + pos_str(s.pos);
      break;
    }
    case NUTDB_SE_NotExpectedKeywords:
      r += "expected keyword (" + join(s.expected_kw) + ") but found token " + s.actual_kw + " at " + pos_str(s.pos);
      break;
    case NUTDB_SE_ParseFail: r += "fail to parse (" + s.msg + ") at " + pos_str(s.pos); break;
    case NUTDB_SE_EmptyQuery: r += "empty query"; break;
    case NUTDB_SE_InvalidEscapedUnicode: r += "invalid escaped unicode '\\u{" + s.raw + "}' in string literal"; break;
    case NUTDB_SE_InvalidFloatLiteral: r += "invalid float '" + s.raw + "'"; break;
    case NUTDB_SE_InvalidHexLiteral: r += "invalid hex '0x" + s.raw + "'"; break;
    case NUTDB_SE_InvalidIntegerLiteral: r += "invalid integer '" + s.raw + "'"; break;
    case NUTDB_SE_Conflicts: r += "(" + s.this_ + ") conflicts with (" + s.that + ") near " + pos_str(s.pos); break;
  }
  return r;
}

// ------------------------------------------------------------------------------------------
// tree -> flat nodes (post-order, include/nutdb_gpu.h)
// ------------------------------------------------------------------------------------------
struct Flat {
  std::vector<NutdbNode> n;
  const char* base = nullptr;
  size_t m_alg = 0;  // Expr + Query + DataType values + 1 Statement (SURVEY.md 8d "M")

  uint32_t off(const char* p) const { return (uint32_t)(p - base); }
  uint32_t leaf(uint8_t kind, uint8_t sub, uint16_t aux, uint32_t a, uint32_t b) {
    n.push_back(NutdbNode{kind, sub, aux, NUTDB_NO_PARENT, a, b});
    return (uint32_t)n.size() - 1;
  }
  uint32_t leaf_sv(uint8_t kind, sv s, uint8_t sub = 0, uint16_t aux = 0) {
    return leaf(kind, sub, aux, off(s.data()), off(s.data()) + (uint32_t)s.size());
  }
  uint32_t leaf_span(uint8_t kind, Span s, uint8_t sub = 0, uint16_t aux = 0) {
    return leaf(kind, sub, aux, (uint32_t)s.start, (uint32_t)s.end);
  }
  size_t mark() const { return n.size(); }
  static bool is_interior(uint8_t kind) { return kind >= NUTDB_NK_FIRST_INTERIOR; }
  uint32_t subtree_start(uint32_t i) const {
    if (is_interior(n[i].kind)) return n[i].a;
    if (n[i].kind == NUTDB_NK_IDENT && (n[i].aux & 1)) return i - 1;
    return i;
  }
  // close an interior node whose children are the subtrees in [m, size)
  uint32_t close(size_t m, uint8_t kind, uint8_t sub = 0, uint16_t aux = 0) {
    uint32_t self = (uint32_t)n.size();
    uint32_t cnt = 0;
    int64_t r = (int64_t)self - 1;
    while (r >= (int64_t)m) {
      n[(size_t)r].parent = self;
      cnt++;
      r = (int64_t)subtree_start((uint32_t)r) - 1;
    }
    n.push_back(NutdbNode{kind, sub, aux, NUTDB_NO_PARENT, (uint32_t)m, cnt});
    return self;
  }

  void expr(const Expr& e) {
    m_alg++;
    switch (e.k) {
      case Expr::Identifier:
        if (e.has_qual) {
          uint32_t q = leaf_span(NUTDB_NK_QUAL, e.qspan);
          uint32_t i = leaf_span(NUTDB_NK_IDENT, e.span, e.sub, 1);
          n[q].parent = i;
        } else {
          leaf_span(NUTDB_NK_IDENT, e.span, e.sub, 0);
        }
        break;
      case Expr::QueryParameter: leaf_span(NUTDB_NK_QPARAM, e.span, 0, e.hex ? 1 : 0); break;
      case Expr::Literal:
        switch (e.sub) {
          case Expr::LInteger: leaf_span(NUTDB_NK_LIT_INT, e.span, e.flag ? 0 : 1, e.hex ? 1 : 0); break;
          case Expr::LFloat: leaf_span(NUTDB_NK_LIT_FLOAT, e.span, e.flag ? 1 : 0, 0); break;
          case Expr::LString: leaf_span(NUTDB_NK_LIT_STR, e.span, e.strkind, 0); break;
          case Expr::LBoolean: leaf(NUTDB_NK_LIT_BOOL, e.flag ? 1 : 0, 0, 0, 0); break;
          case Expr::LInterval: leaf_span(NUTDB_NK_LIT_INTERVAL, e.span, e.unit, e.hex ? 1 : 0); break;
          default: leaf(NUTDB_NK_LIT_NULL, 0, 0, 0, 0);
        }
        break;
      case Expr::Collection: {
        size_t m = mark();
        for (const Expr& k : e.kids) expr(k);
        close(m, NUTDB_NK_COLLECTION, e.sub);
        break;
      }
      case Expr::UnaryOp: {
        size_t m = mark();
        expr(e.kids[0]);
        close(m, NUTDB_NK_UNARY, e.sub);
        break;
      }
      case Expr::BinaryOp: {
        size_t m = mark();
        expr(e.kids[0]);
        expr(e.kids[1]);
        close(m, NUTDB_NK_BINARY, e.sub);
        break;
      }
      case Expr::FnCall: fncall(e); break;
      case Expr::Subquery: query(*e.q); break;
    }
  }
  void fncall(const Expr& e) {
    size_t m = mark();
    if (e.sub == F_Others) leaf_sv(NUTDB_NK_FN_NAME, e.s1);
    for (const Expr& k : e.kids) expr(k);
    close(m, NUTDB_NK_FNCALL, e.sub);
  }
  void query_expr(const QueryExpr& q) {
    expr(q.inner);
    if (q.alias) leaf_sv(NUTDB_NK_ALIAS, *q.alias);
  }
  void source(const QuerySource& s) {
    if (s.kind == QuerySource::TableFn) {
      m_alg++;  // the FnCall was an Expr value before conversion; count it once
      m_alg--;
      fncall_counted(s.fn);
    } else if (s.kind == QuerySource::Table) {
      leaf_sv(NUTDB_NK_IDENT, s.table, 0, 0);
    } else {
      query(*s.q);
    }
    if (s.alias) leaf_sv(NUTDB_NK_ALIAS, *s.alias);
  }
  void fncall_counted(const Expr& e) {  // FnCall stored as a bare FnCall struct (not an Expr) in the reference
    size_t m = mark();
    if (e.sub == F_Others) leaf_sv(NUTDB_NK_FN_NAME, e.s1);
    for (const Expr& k : e.kids) expr(k);
    close(m, NUTDB_NK_FNCALL, e.sub);
  }
  void num(Span s, bool hex) { leaf_span(NUTDB_NK_NUM, s, 0, hex ? 1 : 0); }
  void strlit(const StringLit& s) { leaf_span(NUTDB_NK_STR, s.span, s.strkind, 0); }
  void body(const QueryBody& b) {
    size_t m = mark();
    if (b.with) {
      size_t w = mark();
      for (const QueryCTE& c : *b.with) {
        leaf_sv(NUTDB_NK_NAME, c.alias);
        m_alg++;  // Expr::Subquery wrapper the CTE was parsed through
        m_alg--;
        query(*c.subquery);
      }
      close(w, NUTDB_NK_WITH);
    }
    if (b.has_distinct) {
      size_t d = mark();
      if (b.distinct_on)
        for (const QueryExpr& q : *b.distinct_on) query_expr(q);
      close(d, NUTDB_NK_DISTINCT, 0, b.distinct_on ? 1 : 0);
    }
    {
      size_t c = mark();
      for (const QueryExpr& q : b.columns) query_expr(q);
      close(c, NUTDB_NK_COLS);
    }
    if (b.from) {
      size_t f = mark();
      source(*b.from);
      close(f, NUTDB_NK_FROM);
    }
    for (const JoinClause& j : b.joins) {
      size_t jm = mark();
      source(j.source);
      if (j.is_using) {
        for (const Expr& c : j.using_cols) {
          m_alg--;  // Identifier structs inside Using are not Expr values
          expr(c);
        }
      } else {
        expr(*j.on);
      }
      close(jm, NUTDB_NK_JOIN, (uint8_t)j.typ, j.is_using ? 1 : 0);
    }
    if (b.where) {
      size_t w = mark();
      expr(*b.where);
      close(w, NUTDB_NK_WHERE);
    }
    if (b.group_by) {
      size_t g = mark();
      for (const QueryExpr& q : *b.group_by) query_expr(q);
      close(g, NUTDB_NK_GROUPBY);
    }
    if (b.having) {
      size_t h = mark();
      expr(*b.having);
      close(h, NUTDB_NK_HAVING);
    }
    if (b.order_by) {
      size_t ob = mark();
      for (const OrderKey& k : *b.order_by) {
        query_expr(k.expr);
        if (k.desc) leaf(NUTDB_NK_ORDER_DESC, 0, 0, 0, 0);
      }
      close(ob, NUTDB_NK_ORDERBY);
    }
    if (b.limit) {
      size_t l = mark();
      num(b.limit->s1, b.limit->hex1);
      if (b.limit->form != 0) num(b.limit->s2, b.limit->hex2);
      close(l, NUTDB_NK_LIMIT, (uint8_t)b.limit->form, b.limit->with_ties ? 1 : 0);
    }
    close(m, NUTDB_NK_QUERY_BODY);
  }
  void query(const Query& q) {
    m_alg++;
    if (!q.is_union) {
      body(*q.body);
    } else {
      size_t m = mark();
      query(*q.left);
      query(*q.right);
      close(m, NUTDB_NK_QUERY_UNION, (uint8_t)q.typ);
    }
  }
  void datatype(const DataType& d) {
    m_alg++;
    if (!d.compound) {
      if (d.has_param) {
        size_t m = mark();
        num(d.pspan, d.phex);
        close(m, NUTDB_NK_DT_PARAM, (uint8_t)d.id);
      } else {
        leaf(NUTDB_NK_DT_SCALAR, (uint8_t)d.id, 0, 0, 0);
      }
      return;
    }
    size_t m = mark();
    if (d.id == 1) {
      for (const EnumBind& b : d.binds) {
        strlit(b.literal);
        if (b.has_id) num(b.idspan, b.idhex);
      }
    } else if (d.id == 3) {
      datatype(d.inner[1]);  // key (source order)
      datatype(d.inner[0]);  // value
    } else {
      for (const DataType& i : d.inner) datatype(i);
    }
    close(m, NUTDB_NK_DT_COMPOUND, (uint8_t)d.id);
  }
  void exprs_attr(const std::vector<Expr>& v, uint8_t kind) {
    size_t m = mark();
    for (const Expr& e : v) expr(e);
    close(m, kind);
  }
  void coldef(const ColumnDefinition& c) {
    size_t m = mark();
    leaf_sv(NUTDB_NK_NAME, c.name);
    datatype(c.typ);
    for (int a : c.attr_order) {
      if (a == 0) {
        size_t d = mark();
        expr(*c.default_);
        close(d, NUTDB_NK_ATTR_DEFAULT);
      } else {
        strlit(*c.comment);
      }
    }
    close(m, NUTDB_NK_COLDEF);
  }
  void idxdef(const IndexDefinition& d) {
    size_t m = mark();
    leaf_sv(NUTDB_NK_NAME, d.name);
    fncall_counted(d.indexer);
    close(m, NUTDB_NK_INDEXDEF);
  }
  void condef(const ConstraintDefinition& c) {
    size_t m = mark();
    leaf_sv(NUTDB_NK_NAME, c.name);
    expr(c.check);
    close(m, NUTDB_NK_CONSTRDEF);
  }
  void statement(const Statement& s) {
    m_alg++;
    size_t m = mark();
    switch (s.k) {
      case Statement::Select:
        query(*s.query);
        close(m, NUTDB_NK_STMT_SELECT);
        break;
      case Statement::Explain:
        query(*s.query);
        close(m, NUTDB_NK_STMT_EXPLAIN);
        break;
      case Statement::Insert:
        leaf_sv(NUTDB_NK_NAME, s.table_name);
        if (s.column_list)
          for (sv c : *s.column_list) leaf_sv(NUTDB_NK_NAME, c);
        if (s.insert_kind == 0) {
          size_t rm = mark();
          for (size_t i = 0; i < s.rows_data.size(); i += s.column_size) {
            size_t r = mark();
            for (size_t j = 0; j < s.column_size; j++) expr(s.rows_data[i + j]);
            close(r, NUTDB_NK_ROW);
          }
          close(rm, NUTDB_NK_ROWS);
        } else if (s.insert_kind == 1) {
          query(*s.query);
        } else {
          fncall_counted(s.insert_fn);
        }
        close(m, NUTDB_NK_STMT_INSERT);
        break;
      case Statement::Alter:
        leaf_sv(NUTDB_NK_NAME, s.table_name);
        if (s.alter_action == 0) {
          if (s.entity_kind == 0) coldef(s.col);
          else if (s.entity_kind == 1) idxdef(s.idx);
          else condef(s.con);
          if (s.position == 0) leaf(NUTDB_NK_POS_FIRST, 0, 0, 0, 0);
          else if (s.position == 1) leaf_sv(NUTDB_NK_POS_AFTER, s.after_name);
        } else if (s.alter_action == 1) {
          if (s.entity_kind == 3) strlit(s.partition);
          else leaf_sv(NUTDB_NK_ENT_NAME, s.entity_name, (uint8_t)s.entity_kind);
        } else {
          if (s.entity_kind == 3) leaf(NUTDB_NK_ENT_NAME, 3, 0, 0, 0);
          else leaf_sv(NUTDB_NK_ENT_NAME, s.entity_name, (uint8_t)s.entity_kind);
          leaf_sv(NUTDB_NK_NAME, s.new_name);
        }
        close(m, NUTDB_NK_STMT_ALTER, (uint8_t)s.alter_action, s.flag ? 1 : 0);
        break;
      case Statement::Create: {
        size_t d = mark();
        if (!s.is_view) {
          const TableDefinition& t = s.table;
          leaf_sv(NUTDB_NK_NAME, t.name);
          for (auto& it : t.item_order) {
            if (it.first == 0) coldef(t.columns[it.second]);
            else if (it.first == 1) idxdef(t.indexes[it.second]);
            else condef(t.constraints[it.second]);
          }
          for (int a : t.attr_order) {
            if (a == 0) exprs_attr(*t.primary_key, NUTDB_NK_ATTR_PK);
            else if (a == 1) exprs_attr(*t.order_by, NUTDB_NK_ATTR_ORDER);
            else if (a == 2) { size_t p = mark(); expr(*t.partition_by); close(p, NUTDB_NK_ATTR_PART); }
            else strlit(*t.comment);
          }
          close(d, NUTDB_NK_TABLEDEF);
        } else {
          const ViewDefinition& v = s.view;
          leaf_sv(NUTDB_NK_NAME, v.name);
          for (int a : v.attr_order) {
            if (a == 0) leaf_sv(NUTDB_NK_STRATEGY, v.strategy);
            else if (a == 1) exprs_attr(*v.primary_key, NUTDB_NK_ATTR_PK);
            else if (a == 2) exprs_attr(*v.order_by, NUTDB_NK_ATTR_ORDER);
            else if (a == 3) { size_t p = mark(); expr(*v.partition_by); close(p, NUTDB_NK_ATTR_PART); }
            else strlit(*v.comment);
          }
          query(*v.query);
          close(d, NUTDB_NK_VIEWDEF);
        }
        close(m, NUTDB_NK_STMT_CREATE, 0, s.flag ? 1 : 0);
        break;
      }
      case Statement::Describe:
        if (s.entity != 2) leaf_sv(NUTDB_NK_NAME, s.name);
        close(m, NUTDB_NK_STMT_DESCRIBE, (uint8_t)s.entity);
        break;
      case Statement::Drop:
      case Statement::Truncate:
        leaf_sv(NUTDB_NK_NAME, s.name);
        close(m, s.k == Statement::Drop ? NUTDB_NK_STMT_DROP : NUTDB_NK_STMT_TRUNCATE, (uint8_t)s.entity, s.flag ? 1 : 0);
        break;
      case Statement::Optimize:
        leaf_sv(NUTDB_NK_NAME, s.table_name);
        if (s.partition_key) expr(*s.partition_key);
        close(m, NUTDB_NK_STMT_OPTIMIZE);
        break;
      case Statement::Set:
        leaf_sv(NUTDB_NK_NAME, s.config_name);
        expr(s.value);
        close(m, NUTDB_NK_STMT_SET);
        break;
    }
  }
};

}  // namespace ora
