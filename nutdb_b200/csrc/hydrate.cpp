// hydrate.cpp -- host-side re-hydration of the flat result of nutdb_gpu_parse_batch().
//
// nutdb_fmt_debug renders statement i exactly as Rust's `format!("{:?}", stmt)` would render the
// reference's `Statement` (derive(Debug) on src/parser/ast/{mod,expr,item,query,alter}.rs), and
// nutdb_fmt_error renders its `ParseError` as `Display` does (src/parser/error.rs:8-57,
// tokenizer/error.rs:14-30).  This is the proof that the flat arrays carry the whole AST: a Rust
// re-hydrator into `nutdb::parser::Statement` walks the nodes the same way.
// Pure CPU formatting of results -- nothing here lexes or parses SQL.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nutdb_gpu.h"
#include "lex_tables.hpp"  // KEYWORD_TEXT

namespace hyd {
#include "parse_program.h"  // PARSE_TABLES: keyword lists and expected-token lists of the error sites
}

namespace {

typedef unsigned __int128 u128;

const char* const TT_NAMES[] = {
    "KeywordOrIdentifier", "DelimitedIdentifier", "ConfigIdentifier", "QueryParameter", "RawStringLiteral",
    "EscapedSingleQuotedStringLiteral", "EscapedDoubleQuotedStringLiteral", "IntegerLiteral", "FloatLiteral",
    "HexLiteral", "Comma", "Dot", "Colon", "SemiColon", "Plus", "Minus", "Mul", "Div", "Mod", "Eq", "NotEq", "Lt",
    "Gt", "LtEq", "GtEq", "LParen", "RParen", "LBracket", "RBracket", "LBrace", "RBrace", "BitAnd", "BitOr",
    "BitXor", "BitNot", "BitLShift", "BitRShift", "Comment", "Whitespace", "EOF", "POISON"};
const char* const UNARY_NAMES[] = {"BitwiseNot", "Not", "IsNull", "IsNotNull"};
const char* const BINARY_NAMES[] = {
    "Plus", "Minus", "Multi", "Div", "Mod", "Gt", "Lt", "GtEq", "LtEq", "Eq", "NotEq", "And", "Or", "Xor", "Like",
    "NotLike", "ILike", "NotILike", "In", "NotIn", "IndexAccess", "BitwiseOr", "BitwiseAnd", "BitwiseXor",
    "BitwiseLeftShift", "BitwiseRightShift"};
const char* const FN_NAMES[] = {"If", "MultiIf", "CaseWhen", "Between", "NotBetween", "Exists", "NotExists"};
const char* const COLL_NAMES[] = {"Tuple", "Map", "Array"};
const char* const UNIT_NAMES[] = {"Second", "Minute", "Hour", "Day", "Month", "Year"};
const char* const JOIN_NAMES[] = {"Inner", "FullOuter", "LeftOuter", "RightOuter", "LeftSemi", "RightSemi",
                                  "LeftAnti", "RightAnti", "AsOf"};
const char* const UNION_NAMES[] = {"UnionAll", "UnionDistinct", "Intersect", "Except"};
const char* const SCALAR_NAMES[] = {
    "Int8", "Int16", "Int32", "Int64", "Int128", "UInt8", "UInt16", "UInt32", "UInt64", "UInt128", "Serial32",
    "Serial64", "Serial128", "USerial32", "USerial64", "USerial128", "Decimal32", "Decimal64", "Float32", "Float64",
    "Boolean", "Chars", "String", "Uuid", "Date", "Datetime"};
const char* const COMPOUND_NAMES[] = {"Array", "Enum", "Tuple", "Map", "Dictionary", "Nullable"};

std::string u128_str(u128 v) {
  if (v == 0) return "0";
  std::string s;
  while (v) {
    s.push_back((char)('0' + (int)(v % 10)));
    v /= 10;
  }
  return std::string(s.rbegin(), s.rend());
}

void push_utf8(std::string& o, uint32_t c) {
  if (c < 0x80) o.push_back((char)c);
  else if (c < 0x800) {
    o.push_back((char)(0xC0 | (c >> 6)));
    o.push_back((char)(0x80 | (c & 0x3F)));
  } else if (c < 0x10000) {
    o.push_back((char)(0xE0 | (c >> 12)));
    o.push_back((char)(0x80 | ((c >> 6) & 0x3F)));
    o.push_back((char)(0x80 | (c & 0x3F)));
  } else {
    o.push_back((char)(0xF0 | (c >> 18)));
    o.push_back((char)(0x80 | ((c >> 12) & 0x3F)));
    o.push_back((char)(0x80 | ((c >> 6) & 0x3F)));
    o.push_back((char)(0x80 | (c & 0x3F)));
  }
}

// <str as Debug>::fmt for the characters SQL text can hold (ASCII control characters escaped,
// everything else passed through)
void dbg_str(std::string& o, const std::string& s) {
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

struct H {
  const NutdbNode* nd;  // nodes of this statement (index 0 = first node)
  uint32_t n;
  const uint8_t* sql;
  size_t len;
  std::string o;

  std::string text(uint32_t a, uint32_t b) const {
    if (a > b || b > len) return std::string();
    return std::string((const char*)sql + a, b - a);
  }
  std::string span(uint32_t i) const { return text(nd[i].a, nd[i].b); }
  static bool interior(const NutdbNode& x) { return x.kind >= NUTDB_NK_FIRST_INTERIOR; }
  uint32_t subtree_start(uint32_t i) const {
    if (interior(nd[i])) return nd[i].a;
    if (nd[i].kind == NUTDB_NK_IDENT && (nd[i].aux & 1)) return i - 1;
    return i;
  }
  std::vector<uint32_t> kids(uint32_t i) const {
    std::vector<uint32_t> r;
    if (!interior(nd[i])) return r;
    int64_t k = (int64_t)i - 1;
    while (k >= (int64_t)nd[i].a) {
      r.push_back((uint32_t)k);
      k = (int64_t)subtree_start((uint32_t)k) - 1;
    }
    return std::vector<uint32_t>(r.rbegin(), r.rend());
  }

  u128 integer(uint32_t a, uint32_t b, bool hex) const {  // digits were validated on the device
    u128 v = 0;
    for (uint32_t p = a; p < b && p < len; p++) {
      uint8_t c = sql[p];
      uint32_t d = c <= '9' ? (uint32_t)(c - '0') : (uint32_t)((c | 0x20) - 'a' + 10);
      v = v * (hex ? 16u : 10u) + d;
    }
    return v;
  }
  u128 node_int(uint32_t i) const { return integer(nd[i].a, nd[i].b, (nd[i].aux & 1) != 0); }

  // BigDecimal::from_str + Display (bigdecimal 0.3: digits + scale, no normalisation)
  std::string decimal(uint32_t i) const {
    std::string s = span(i);
    size_t dot = s.find('.');
    std::string digits = dot == std::string::npos ? s : s.substr(0, dot) + s.substr(dot + 1);
    int64_t scale = dot == std::string::npos ? 0 : (int64_t)(s.size() - dot - 1);
    size_t nz = digits.find_first_not_of('0');
    std::string mag = nz == std::string::npos ? "0" : digits.substr(nz);
    std::string before, after;
    if (scale >= (int64_t)mag.size()) {
      after = std::string((size_t)(scale - (int64_t)mag.size()), '0') + mag;
      before = "0";
    } else {
      size_t loc = (size_t)((int64_t)mag.size() - scale);
      after = mag.substr(loc);
      before = mag.substr(0, loc);
    }
    std::string r = after.empty() ? before : before + "." + after;
    return ((nd[i].sub & 1) && mag != "0") ? "-" + r : r;
  }

  // unescape_{single,double}_quoted_string (literal.rs:45-102); literals were validated on the device
  std::string unescape(uint32_t a, uint32_t b, uint8_t kind) const {
    std::string raw = text(a, b);
    if (kind == 0) return raw;
    const char quote = kind == 1 ? '\'' : '"';
    std::string res;
    size_t p = 0;
    auto next = [&](uint32_t& c) -> bool {  // one UTF-8 char
      if (p >= raw.size()) return false;
      uint8_t x = (uint8_t)raw[p];
      size_t w = x < 0x80 ? 1 : (x >= 0xF0 ? 4 : (x >= 0xE0 ? 3 : 2));
      if (x < 0x80) c = x;
      else if (w == 2) c = ((x & 0x1F) << 6) | ((uint8_t)raw[p + 1] & 0x3F);
      else if (w == 3) c = ((x & 0x0F) << 12) | (((uint8_t)raw[p + 1] & 0x3F) << 6) | ((uint8_t)raw[p + 2] & 0x3F);
      else c = ((x & 0x07) << 18) | (((uint8_t)raw[p + 1] & 0x3F) << 12) | (((uint8_t)raw[p + 2] & 0x3F) << 6) |
               ((uint8_t)raw[p + 3] & 0x3F);
      p += w;
      return true;
    };
    uint32_t c;
    while (next(c)) {
      if (c == (uint32_t)quote) {
        uint32_t skip;
        next(skip);
        res.push_back(quote);
      } else if (c == '\\') {
        uint32_t e;
        if (!next(e)) break;
        if (e == 'n') res.push_back('\n');
        else if (e == 'r') res.push_back('\r');
        else if (e == 't') res.push_back('\t');
        else if (e == 'u') {
          uint32_t b2;
          if (!next(b2)) {
            res.push_back('u');
            break;
          }
          if (b2 != '{') {
            res.push_back('u');  // the char after `u` is dropped (literal.rs:70,89-91)
            continue;
          }
          uint32_t v = 0, h;
          bool first = true;
          while (next(h) && h != '}') {
            if (first && h == '+') {
              first = false;
              continue;
            }
            first = false;
            uint32_t d = h <= '9' ? h - '0' : ((h | 0x20) - 'a' + 10);
            v = v * 16 + d;
          }
          push_utf8(res, v);
        } else {
          push_utf8(res, e);
        }
      } else {
        push_utf8(res, c);
      }
    }
    return res;
  }
  std::string str_value(uint32_t i) const { return unescape(nd[i].a, nd[i].b, nd[i].sub); }

  // ------------------------------------------------------------------ expressions
  void identifier(uint32_t i) {
    o += "Identifier { name: ";
    if (nd[i].sub == 1) o += "Wildcard";
    else {
      o += "Word(";
      dbg_str(o, span(i));
      o += ")";
    }
    o += ", qualifier: ";
    if (nd[i].aux & 1) {
      o += "Some(";
      dbg_str(o, span(i - 1));
      o += ")";
    } else {
      o += "None";
    }
    o += " }";
  }
  void fncall(uint32_t i) {
    std::vector<uint32_t> k = kids(i);
    size_t first = 0;
    o += "FnCall { callee: ";
    if (nd[i].sub == 7) {
      o += "Others(";
      dbg_str(o, span(k[0]));
      o += ")";
      first = 1;
    } else {
      o += FN_NAMES[nd[i].sub];
    }
    o += ", arguments: [";
    for (size_t j = first; j < k.size(); j++) {
      if (j > first) o += ", ";
      expr(k[j]);
    }
    o += "] }";
  }
  void expr(uint32_t i) {
    const NutdbNode& x = nd[i];
    switch (x.kind) {
      case NUTDB_NK_IDENT:
        o += "Identifier(";
        identifier(i);
        o += ")";
        break;
      case NUTDB_NK_QPARAM: o += "QueryParameter(QueryParameter { index: " + u128_str(node_int(i)) + " })"; break;
      case NUTDB_NK_LIT_INT:
        o += "Literal(Integer(" + u128_str(node_int(i)) + ", " + ((x.sub & 1) ? "false" : "true") + "))";
        break;
      case NUTDB_NK_LIT_FLOAT: o += "Literal(Float(BigDecimal(\"" + decimal(i) + "\")))"; break;
      case NUTDB_NK_LIT_STR:
        o += "Literal(String(";
        dbg_str(o, str_value(i));
        o += "))";
        break;
      case NUTDB_NK_LIT_BOOL: o += x.sub ? "Literal(Boolean(true))" : "Literal(Boolean(false))"; break;
      case NUTDB_NK_LIT_NULL: o += "Literal(Null)"; break;
      case NUTDB_NK_LIT_INTERVAL:
        o += "Literal(Interval(" + u128_str(node_int(i)) + ", " + UNIT_NAMES[x.sub] + "))";
        break;
      case NUTDB_NK_COLLECTION: {
        o += std::string("Collection(Collection { typ: ") + COLL_NAMES[x.sub] + ", items: [";
        std::vector<uint32_t> k = kids(i);
        for (size_t j = 0; j < k.size(); j++) {
          if (j) o += ", ";
          expr(k[j]);
        }
        o += "] })";
        break;
      }
      case NUTDB_NK_UNARY:
        o += std::string("UnaryOp(UnaryOp { op: ") + UNARY_NAMES[x.sub] + ", operand: ";
        expr(kids(i)[0]);
        o += " })";
        break;
      case NUTDB_NK_BINARY: {
        std::vector<uint32_t> k = kids(i);
        o += std::string("BinaryOp(BinaryOp { op: ") + BINARY_NAMES[x.sub] + ", left: ";
        expr(k[0]);
        o += ", right: ";
        expr(k[1]);
        o += " })";
        break;
      }
      case NUTDB_NK_FNCALL:
        o += "FnCall(";
        fncall(i);
        o += ")";
        break;
      case NUTDB_NK_QUERY_BODY:
      case NUTDB_NK_QUERY_UNION:
        o += "Subquery(";
        query(i);
        o += ")";
        break;
      default: o += "?"; break;
    }
  }
  // (expr ALIAS?)* starting at k[j]; consumes one QueryExpr
  void query_expr(const std::vector<uint32_t>& k, size_t& j) {
    o += "QueryExpr { inner: ";
    expr(k[j++]);
    o += ", alias: ";
    if (j < k.size() && nd[k[j]].kind == NUTDB_NK_ALIAS) {
      o += "Some(";
      dbg_str(o, span(k[j++]));
      o += ")";
    } else {
      o += "None";
    }
    o += " }";
  }
  void query_expr_list(uint32_t i) {
    std::vector<uint32_t> k = kids(i);
    o += "[";
    for (size_t j = 0; j < k.size();) {
      if (j) o += ", ";
      query_expr(k, j);
    }
    o += "]";
  }
  void source(const std::vector<uint32_t>& k, size_t& j) {
    o += "QuerySource { inner: ";
    uint32_t s = k[j++];
    if (nd[s].kind == NUTDB_NK_FNCALL) {
      o += "TableFn(";
      fncall(s);
      o += ")";
    } else if (nd[s].kind == NUTDB_NK_IDENT) {
      o += "Table(";
      dbg_str(o, span(s));
      o += ")";
    } else {
      o += "Subquery(";
      query(s);
      o += ")";
    }
    o += ", alias: ";
    if (j < k.size() && nd[k[j]].kind == NUTDB_NK_ALIAS) {
      o += "Some(";
      dbg_str(o, span(k[j++]));
      o += ")";
    } else {
      o += "None";
    }
    o += " }";
  }
  void query(uint32_t i) {
    if (nd[i].kind == NUTDB_NK_QUERY_BODY) {
      o += "Single(";
      body(i);
      o += ")";
    } else {
      std::vector<uint32_t> k = kids(i);
      o += std::string("Union { typ: ") + UNION_NAMES[nd[i].sub] + ", left: ";
      query(k[0]);
      o += ", right: ";
      query(k[1]);
      o += " }";
    }
  }
  void body(uint32_t i) {
    std::vector<uint32_t> k = kids(i);
    size_t j = 0;
    auto is = [&](uint8_t kind) { return j < k.size() && nd[k[j]].kind == kind; };
    o += "QueryBody { with: ";
    if (is(NUTDB_NK_WITH)) {
      std::vector<uint32_t> w = kids(k[j++]);
      o += "Some(WithClause { cte_list: [";
      for (size_t c = 0; c + 1 < w.size(); c += 2) {
        if (c) o += ", ";
        o += "QueryCTE { subquery: ";
        query(w[c + 1]);
        o += ", alias: ";
        dbg_str(o, span(w[c]));
        o += " }";
      }
      o += "] })";
    } else {
      o += "None";
    }
    o += ", distinct: ";
    if (is(NUTDB_NK_DISTINCT)) {
      uint32_t d = k[j++];
      o += "Some(DistinctClause { columns: ";
      if (nd[d].aux & 1) {
        o += "Some(";
        query_expr_list(d);
        o += ")";
      } else {
        o += "None";
      }
      o += " })";
    } else {
      o += "None";
    }
    o += ", columns: ";
    query_expr_list(k[j++]);
    o += ", from: ";
    if (is(NUTDB_NK_FROM)) {
      std::vector<uint32_t> f = kids(k[j++]);
      size_t p = 0;
      o += "Some(FromClause { source: ";
      source(f, p);
      o += " })";
    } else {
      o += "None";
    }
    o += ", joins: [";
    bool firstj = true;
    while (is(NUTDB_NK_JOIN)) {
      uint32_t jn = k[j++];
      std::vector<uint32_t> c = kids(jn);
      size_t p = 0;
      if (!firstj) o += ", ";
      firstj = false;
      o += std::string("JoinClause { typ: ") + JOIN_NAMES[nd[jn].sub] + ", source: ";
      source(c, p);
      o += ", condition: ";
      if (nd[jn].aux & 1) {
        o += "Using([";
        for (size_t q = p; q < c.size(); q++) {
          if (q > p) o += ", ";
          identifier(c[q]);
        }
        o += "])";
      } else {
        o += "On(";
        expr(c[p]);
        o += ")";
      }
      o += " }";
    }
    o += "], where: ";
    if (is(NUTDB_NK_WHERE)) {
      o += "Some(WhereClause { condition: ";
      expr(kids(k[j++])[0]);
      o += " })";
    } else {
      o += "None";
    }
    o += ", group_by: ";
    if (is(NUTDB_NK_GROUPBY)) {
      o += "Some(GroupByClause { keys: ";
      query_expr_list(k[j++]);
      o += " })";
    } else {
      o += "None";
    }
    o += ", having: ";
    if (is(NUTDB_NK_HAVING)) {
      o += "Some(HavingClause { condition: ";
      expr(kids(k[j++])[0]);
      o += " })";
    } else {
      o += "None";
    }
    o += ", order_by: ";
    if (is(NUTDB_NK_ORDERBY)) {
      std::vector<uint32_t> c = kids(k[j++]);
      o += "Some(OrderByClause { keys: [";
      for (size_t p = 0; p < c.size();) {
        if (p) o += ", ";
        o += "QueryOrderKey { expr: ";
        query_expr(c, p);
        o += ", direction: ";
        if (p < c.size() && nd[c[p]].kind == NUTDB_NK_ORDER_DESC) {
          o += "DESC";
          p++;
        } else {
          o += "ASC";
        }
        o += " }";
      }
      o += "] })";
    } else {
      o += "None";
    }
    o += ", limit: ";
    if (is(NUTDB_NK_LIMIT)) {
      uint32_t l = k[j++];
      std::vector<uint32_t> c = kids(l);
      u128 first = node_int(c[0]), second = c.size() > 1 ? node_int(c[1]) : 0;
      u128 size = nd[l].sub == 1 ? second : first, offset = nd[l].sub == 0 ? 0 : (nd[l].sub == 1 ? first : second);
      o += "Some(LimitClause { size: " + u128_str(size) + ", offset: " + u128_str(offset) +
           ", with_ties: " + ((nd[l].aux & 1) ? "true" : "false") + " })";
    } else {
      o += "None";
    }
    o += " }";
  }

  // ------------------------------------------------------------------ DDL
  void datatype(uint32_t i) {
    const NutdbNode& x = nd[i];
    if (x.kind == NUTDB_NK_DT_SCALAR) {
      o += std::string("Scalar(") + SCALAR_NAMES[x.sub];
      if (x.sub == 22) o += " { max_length: 0 }";
      o += ")";
      return;
    }
    std::vector<uint32_t> k = kids(i);
    if (x.kind == NUTDB_NK_DT_PARAM) {
      const char* field = (x.sub == 16 || x.sub == 17) ? "scale" : (x.sub == 21 ? "length" : "max_length");
      o += std::string("Scalar(") + SCALAR_NAMES[x.sub] + " { " + field + ": " + u128_str(node_int(k[0])) + " })";
      return;
    }
    o += std::string("Compound(") + COMPOUND_NAMES[x.sub] + "(";
    if (x.sub == 1) {  // Enum: explicit `= n` resets the counter (mod.rs:1799-1813)
      o += "[";
      u128 id = 0;
      bool first = true;
      for (size_t j = 0; j < k.size();) {
        uint32_t lit = k[j++];
        if (j < k.size() && nd[k[j]].kind == NUTDB_NK_NUM) id = node_int(k[j++]);
        if (!first) o += ", ";
        first = false;
        o += "EnumBind { id: " + u128_str(id) + ", literal: ";
        dbg_str(o, str_value(lit));
        o += " }";
        id += 1;
      }
      o += "]";
    } else if (x.sub == 2) {
      o += "[";
      for (size_t j = 0; j < k.size(); j++) {
        if (j) o += ", ";
        datatype(k[j]);
      }
      o += "]";
    } else if (x.sub == 3) {  // Map(K, V) is stored as Map(Box(V), Box(K)) (mod.rs:1776-1780)
      datatype(k[1]);
      o += ", ";
      datatype(k[0]);
    } else {
      datatype(k[0]);
    }
    o += "))";
  }
  void opt_strlit(int64_t i) {
    if (i < 0) {
      o += "None";
      return;
    }
    o += "Some(";
    dbg_str(o, str_value((uint32_t)i));
    o += ")";
  }
  void expr_list_of(uint32_t i) {
    std::vector<uint32_t> k = kids(i);
    o += "[";
    for (size_t j = 0; j < k.size(); j++) {
      if (j) o += ", ";
      expr(k[j]);
    }
    o += "]";
  }
  void coldef(uint32_t i) {
    std::vector<uint32_t> k = kids(i);
    int64_t def = -1, com = -1;
    for (size_t j = 2; j < k.size(); j++) {
      if (nd[k[j]].kind == NUTDB_NK_ATTR_DEFAULT) def = k[j];
      else com = k[j];
    }
    o += "ColumnDefinition { name: ";
    dbg_str(o, span(k[0]));
    o += ", typ: ";
    datatype(k[1]);
    o += ", default: ";
    if (def >= 0) {
      o += "Some(";
      expr(kids((uint32_t)def)[0]);
      o += ")";
    } else {
      o += "None";
    }
    o += ", comment: ";
    opt_strlit(com);
    o += " }";
  }
  void condef(uint32_t i) {
    std::vector<uint32_t> k = kids(i);
    o += "ConstraintDefinition { name: ";
    dbg_str(o, span(k[0]));
    o += ", check: ";
    expr(k[1]);
    o += " }";
  }
  void idxdef(uint32_t i) {
    std::vector<uint32_t> k = kids(i);
    o += "IndexDefinition { name: ";
    dbg_str(o, span(k[0]));
    o += ", indexer: ";
    fncall(k[1]);
    o += " }";
  }
  // primary_key / order_by / partition_by / comment in field order, wherever they appeared
  void common_attrs(const std::vector<uint32_t>& k, size_t from, size_t to, int64_t& pk, int64_t& ob, int64_t& pb,
                    int64_t& com, int64_t& strat) {
    pk = ob = pb = com = strat = -1;
    for (size_t j = from; j < to; j++) {
      switch (nd[k[j]].kind) {
        case NUTDB_NK_ATTR_PK: pk = k[j]; break;
        case NUTDB_NK_ATTR_ORDER: ob = k[j]; break;
        case NUTDB_NK_ATTR_PART: pb = k[j]; break;
        case NUTDB_NK_STR: com = k[j]; break;
        case NUTDB_NK_STRATEGY: strat = k[j]; break;
        default: break;
      }
    }
  }
  void opt_list_attr(int64_t i) {
    if (i < 0) {
      o += "None";
      return;
    }
    o += "Some(";
    expr_list_of((uint32_t)i);
    o += ")";
  }
  void opt_expr_attr(int64_t i) {
    if (i < 0) {
      o += "None";
      return;
    }
    o += "Some(";
    expr(kids((uint32_t)i)[0]);
    o += ")";
  }

  void statement(uint32_t i) {
    const NutdbNode& x = nd[i];
    std::vector<uint32_t> k = kids(i);
    switch (x.kind) {
      case NUTDB_NK_STMT_SELECT:
        o += "Select(SelectStmt { query: ";
        query(k[0]);
        o += " })";
        break;
      case NUTDB_NK_STMT_EXPLAIN:
        o += "Explain(ExplainStmt { query: ";
        query(k[0]);
        o += " })";
        break;
      case NUTDB_NK_STMT_INSERT: {
        o += "Insert(InsertStmt { table_name: ";
        dbg_str(o, span(k[0]));
        o += ", column_list: ";
        size_t last = k.size() - 1;
        if (last > 1) {
          o += "Some([";
          for (size_t j = 1; j < last; j++) {
            if (j > 1) o += ", ";
            dbg_str(o, span(k[j]));
          }
          o += "])";
        } else {
          o += "None";
        }
        o += ", data: ";
        uint32_t d = k[last];
        if (nd[d].kind == NUTDB_NK_ROWS) {
          std::vector<uint32_t> rows = kids(d);
          o += "Rows { column_size: " + std::to_string(nd[rows[0]].b) + ", data: [";
          bool first = true;
          for (uint32_t r : rows)
            for (uint32_t e : kids(r)) {
              if (!first) o += ", ";
              first = false;
              expr(e);
            }
          o += "] }";
        } else if (nd[d].kind == NUTDB_NK_FNCALL) {
          o += "FnCall(";
          fncall(d);
          o += ")";
        } else {
          o += "Subquery(";
          query(d);
          o += ")";
        }
        o += " })";
        break;
      }
      case NUTDB_NK_STMT_ALTER: {
        o += "Alter(AlterStmt { alter: Alter { action: ";
        const char* flag = (x.aux & 1) ? "true" : "false";
        if (x.sub == 0) {
          uint32_t e = k[1];
          o += "Add { entity: ";
          if (nd[e].kind == NUTDB_NK_COLDEF) { o += "Column("; coldef(e); o += ")"; }
          else if (nd[e].kind == NUTDB_NK_INDEXDEF) { o += "Index("; idxdef(e); o += ")"; }
          else { o += "Constraint("; condef(e); o += ")"; }
          o += std::string(", if_not_exists: ") + flag + ", position: ";
          if (k.size() > 2 && nd[k[2]].kind == NUTDB_NK_POS_FIRST) o += "First";
          else if (k.size() > 2 && nd[k[2]].kind == NUTDB_NK_POS_AFTER) { o += "After("; dbg_str(o, span(k[2])); o += ")"; }
          else o += "Last";
          o += " }";
        } else if (x.sub == 1) {
          uint32_t e = k[1];
          static const char* const names[] = {"Column", "Index", "Constraint"};
          o += "Drop { entity: ";
          if (nd[e].kind == NUTDB_NK_STR) { o += "Partition("; dbg_str(o, str_value(e)); }
          else { o += names[nd[e].sub]; o += "("; dbg_str(o, span(e)); }
          o += std::string("), if_exists: ") + flag + " }";
        } else {
          uint32_t e = k[1];
          static const char* const names[] = {"Column", "Index", "Constraint", "Table"};
          o += "Rename { entity: ";
          o += names[nd[e].sub];
          if (nd[e].sub != 3) { o += "("; dbg_str(o, span(e)); o += ")"; }
          o += ", new_name: ";
          dbg_str(o, span(k[2]));
          o += " }";
        }
        o += ", table_name: ";
        dbg_str(o, span(k[0]));
        o += " } })";
        break;
      }
      case NUTDB_NK_STMT_CREATE: {
        uint32_t d = k[0];
        std::vector<uint32_t> c = kids(d);
        int64_t pk, ob, pb, com, strat;
        o += std::string("Create(CreateStmt { if_not_exists: ") + ((x.aux & 1) ? "true" : "false") + ", entity_def: ";
        if (nd[d].kind == NUTDB_NK_TABLEDEF) {
          o += "Table(TableDefinition { name: ";
          dbg_str(o, span(c[0]));
          const uint8_t item_kinds[3] = {NUTDB_NK_COLDEF, NUTDB_NK_CONSTRDEF, NUTDB_NK_INDEXDEF};
          const char* const labels[3] = {", columns: [", "], constraints: [", "], indexes: ["};
          for (int g = 0; g < 3; g++) {
            o += labels[g];
            bool first = true;
            for (size_t j = 1; j < c.size(); j++) {
              if (nd[c[j]].kind != item_kinds[g]) continue;
              if (!first) o += ", ";
              first = false;
              if (g == 0) coldef(c[j]);
              else if (g == 1) condef(c[j]);
              else idxdef(c[j]);
            }
          }
          common_attrs(c, 1, c.size(), pk, ob, pb, com, strat);
          o += "], primary_key: ";
          opt_list_attr(pk);
          o += ", order_by: ";
          opt_list_attr(ob);
          o += ", partition_by: ";
          opt_expr_attr(pb);
          o += ", comment: ";
          opt_strlit(com);
          o += " })";
        } else {
          common_attrs(c, 1, c.size() - 1, pk, ob, pb, com, strat);
          o += "View(ViewDefinition { name: ";
          dbg_str(o, span(c[0]));
          o += ", strategy: ";
          dbg_str(o, strat >= 0 ? span((uint32_t)strat) : std::string());
          o += ", primary_key: ";
          opt_list_attr(pk);
          o += ", order_by: ";
          opt_list_attr(ob);
          o += ", partition_by: ";
          opt_expr_attr(pb);
          o += ", query: ";
          query(c.back());
          o += ", comment: ";
          opt_strlit(com);
          o += " })";
        }
        o += " })";
        break;
      }
      case NUTDB_NK_STMT_DESCRIBE:
        o += "Describe(DescribeStmt { entity: ";
        if (x.sub == 2) o += "Database";
        else {
          o += x.sub == 0 ? "Table(" : "View(";
          dbg_str(o, span(k[0]));
          o += ")";
        }
        o += " })";
        break;
      case NUTDB_NK_STMT_DROP:
      case NUTDB_NK_STMT_TRUNCATE:
        o += x.kind == NUTDB_NK_STMT_DROP ? "Drop(DropStmt { typ: " : "Truncate(TruncateStmt { typ: ";
        o += x.sub == 0 ? "Table" : "View";
        o += std::string(", if_exists: ") + ((x.aux & 1) ? "true" : "false") + ", name: ";
        dbg_str(o, span(k[0]));
        o += " })";
        break;
      case NUTDB_NK_STMT_OPTIMIZE:
        o += "Optimize(OptimizeStmt { table_name: ";
        dbg_str(o, span(k[0]));
        o += ", partition_key: ";
        if (k.size() > 1) {
          o += "Some(";
          expr(k[1]);
          o += ")";
        } else {
          o += "None";
        }
        o += " })";
        break;
      case NUTDB_NK_STMT_SET:
        o += "Set(SetStmt { config_name: ";
        dbg_str(o, span(k[0]));
        o += ", value: ";
        expr(k[1]);
        o += " })";
        break;
      default: o += "?"; break;
    }
  }
};

size_t deliver(const std::string& s, char* buf, size_t cap) {
  if (buf && cap) {
    size_t m = s.size() < cap - 1 ? s.size() : cap - 1;
    std::memcpy(buf, s.data(), m);
    buf[m] = 0;
  }
  return s.size() + 1;
}

std::string pos_str(const NutdbError& e) { return "line " + std::to_string(e.line) + " col " + std::to_string(e.col); }

std::string char_at(const uint8_t* sql, size_t len, uint32_t pos) {
  if (pos >= len) return std::string();
  uint8_t x = sql[pos];
  size_t w = x < 0x80 ? 1 : (x >= 0xF0 ? 4 : (x >= 0xE0 ? 3 : 2));
  if (pos + w > len) w = len - pos;
  return std::string((const char*)sql + pos, w);
}

std::string lex_error_text(const NutdbError& e, const uint8_t* sql, size_t len) {
  std::string c = "'" + char_at(sql, len, e.pos) + "'";
  const char* kind = "Unexpected Char";
  std::string ctx;
  switch (e.code) {
    case NUTDB_LE_INVALID_CHAR: ctx = c + " is invalid outside string literal"; break;
    case NUTDB_LE_STR_CR: ctx = "\\r in string is supported but should be escaped by '\\'"; break;
    case NUTDB_LE_STR_LF: ctx = "\\n in string is supported but should be escaped by '\\'"; break;
    case NUTDB_LE_STR_EOF: kind = "Unexpected EOF"; ctx = "string literal is not complete"; break;
    case NUTDB_LE_NUM_ZERO: ctx = c + " is invalid in numeric literal"; break;
    case NUTDB_LE_NUM_INT: ctx = c + " cannot be a part of integer literal"; break;
    case NUTDB_LE_NUM_FLOAT: ctx = c + " cannot be a part of float literal"; break;
    case NUTDB_LE_IDENT_END: ctx = c + " cannot be a part of identifier or keyword"; break;
    case NUTDB_LE_CFG_DIGIT: ctx = "config identifier cannot starts with numbers"; break;
    case NUTDB_LE_CFG_END: ctx = c + " cannot be a part of config identifier"; break;
    case NUTDB_LE_CFG_EMPTY: kind = "Incomplete Token"; ctx = "identifier should have name"; break;
    case NUTDB_LE_BT_EMPTY: kind = "Incomplete Token"; ctx = "delimited identifier cannot be an empty string"; break;
    case NUTDB_LE_BT_NL: ctx = "'\\r' or '\\n' cannot be a part of delimited identifier"; break;
    case NUTDB_LE_BT_EOF: kind = "Unexpected EOF"; ctx = "delimited identifier is not complete"; break;
    case NUTDB_LE_QP_END: ctx = c + " cannot be a part of query parameter"; break;
    case NUTDB_LE_QP_EMPTY: kind = "Incomplete Token"; ctx = "query parameter should have an index"; break;
    case NUTDB_LE_BANG: ctx = "'!' can only be used with '='"; break;
    case NUTDB_LE_BC_EOF: kind = "Unexpected EOF"; ctx = "block comment is not complete"; break;
    default: ctx = "?"; break;
  }
  return std::string("Lex Error: ") + kind + ": " + ctx + " near " + pos_str(e);
}

std::string syntax_error_text(const NutdbError& e, const uint8_t* sql, size_t len) {
  const hyd::ParseTables& P = hyd::PARSE_TABLES;
  std::string r = "Syntax Error: ";
  auto raw = [&]() { return (e.b <= e.c && e.c <= len) ? std::string((const char*)sql + e.b, e.c - e.b) : std::string(); };
  switch (e.code) {
    case NUTDB_SE_NotExpectedTokenTypes: {
      std::string names;
      if (e.a < NUTDB_EL_COUNT)
        for (uint32_t i = 0; i < P.expected_len[e.a]; i++) {
          if (i) names += ", ";
          names += TT_NAMES[P.expected_list[e.a][i]];
        }
      r += "expected token (" + names + ") but found token " + (e.b <= NUTDB_TT_POISON ? TT_NAMES[e.b] : "?") + " at " +
           pos_str(e);
      break;
    }
    case NUTDB_SE_NotExpectedKeywords: {
      std::string names, actual = raw();
      if (e.a >= NUTDB_KL_SINGLE) {
        uint32_t kw = e.a - NUTDB_KL_SINGLE;
        if (kw >= 1 && kw <= NUTDB_KW_COUNT) names = nlex::KEYWORD_TEXT[kw - 1];
      } else if (e.a < NUTDB_KL_COUNT) {
        for (uint32_t i = 0; i < P.kwlist_len[e.a]; i++) {
          if (i) names += ", ";
          names += nlex::KEYWORD_TEXT[P.kwlist[P.kwlist_off[e.a] + i] - 1];
        }
        if (e.a == NUTDB_KL_VIEW_NEEDS_UPDATE) actual = "as";  // mod.rs:826 reports the constant AS
      }
      r += "expected keyword (" + names + ") but found token " + actual + " at " + pos_str(e);
      break;
    }
    case NUTDB_SE_ParseFail: {
      static const char* const msgs[] = {"?", "statements should start with a keyword", "more than one statement",
                                         "cannot recognize statement", "not a subquery",
                                         "query source must be a subquery, a table function or a table",
                                         "insert source must be a subquery, values, or a function call",
                                         "indexer must be a function call", "`not exists` should have arguments",
                                         "`exists` should have arguments"};
      r += std::string("fail to parse (") + (e.a < 10 ? msgs[e.a] : "?") + ") at " + pos_str(e);
      break;
    }
    case NUTDB_SE_EmptyQuery: r += "empty query"; break;
    case NUTDB_SE_InvalidEscapedUnicode: r += "invalid escaped unicode '\\u{" + raw() + "}' in string literal"; break;
    case NUTDB_SE_InvalidFloatLiteral: r += "invalid float '" + raw() + "'"; break;
    case NUTDB_SE_InvalidHexLiteral: r += "invalid hex '0x" + raw() + "'"; break;
    case NUTDB_SE_InvalidIntegerLiteral: r += "invalid integer '" + raw() + "'"; break;
    case NUTDB_SE_Conflicts: {
      static const char* const what[] = {"?", "primary key", "order by", "partition by", "comment", "update by", "default"};
      std::string a, b;
      if (e.a == NUTDB_CF_ROW_WIDTH) {
        a = "row has " + std::to_string(e.b) + " column(s)";
        b = "previous rows have " + std::to_string(e.c) + " column(s)";
      } else {
        a = b = e.a < 7 ? what[e.a] : "?";
      }
      r += "(" + a + ") conflicts with (" + b + ") near " + pos_str(e);
      break;
    }
    default: r += "?"; break;
  }
  return r;
}

const NutdbError* find_error(const NutdbBatch* b, uint64_t i) {
  uint64_t lo = 0, hi = b->n_err;  // sorted by .stmt
  while (lo < hi) {
    uint64_t mid = (lo + hi) / 2;
    if (b->err[mid].stmt < i) lo = mid + 1;
    else hi = mid;
  }
  return (lo < b->n_err && b->err[lo].stmt == i) ? &b->err[lo] : nullptr;
}

}  // namespace

// Wire nodes (32-bit words, NUTDB_PN_*) of one statement -> NutdbNode records: spans from the running position and
// the (gap, length) fields, subtree starts from the sizes, child counts and parent links by walking each interior
// node's children backwards from the node before it (post-order: a child's subtree ends right before its next
// sibling's begins).
static const NutdbNodeExt* find_ext(const NutdbBatch* b, uint32_t node_index) {
  uint64_t lo = 0, hi = b->n_ext;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) / 2;
    if (b->ext[mid].index < node_index) lo = mid + 1;
    else hi = mid;
  }
  return (lo < b->n_ext && b->ext[lo].index == node_index) ? &b->ext[lo] : nullptr;
}
// false: a malformed stream (an escape without its side-table entry, a subtree that starts before the statement)
static bool expand_stmt_nodes(const NutdbBatch* b, uint64_t begin, uint32_t count, NutdbNode* out) {
  const uint32_t* w = b->pnode + begin;
  uint32_t pos = 0;
  // pass 1: kind / sub / aux, spans (leaves), subtree starts (interior nodes: kept in .a)
  for (uint32_t j = 0; j < count; j++) {
    NutdbNode& o = out[j];
    const uint32_t x = w[j];
    o.kind = (uint8_t)(x & 127u);
    o.sub = (uint8_t)((x >> NUTDB_PN_SUB_SHIFT) & 31u);
    o.aux = (uint16_t)((x >> NUTDB_PN_FLAG_SHIFT) & 1u);
    o.parent = NUTDB_NO_PARENT;
    if (o.kind >= NUTDB_NK_FIRST_INTERIOR) {
      const uint32_t size = x >> NUTDB_PN_SIZE_SHIFT;
      if (size == NUTDB_PN_SIZE_EXT) {
        const NutdbNodeExt* e = find_ext(b, (uint32_t)(begin + j));
        if (!e) return false;
        o.kind = (uint8_t)(e->hdr & 255u);
        o.sub = (uint8_t)((e->hdr >> 8) & 255u);
        o.aux = (uint16_t)(e->hdr >> 16);
        o.a = e->a;
      } else {
        if (size > j) return false;
        o.a = j - size;
      }
      o.b = 0;
    } else {
      const uint32_t gap = (x >> NUTDB_PN_GAP_SHIFT) & 1023u, len = x >> NUTDB_PN_LEN_SHIFT;
      if (len == NUTDB_PN_LEN_SPECIAL && gap == NUTDB_PN_GAP_NOSPAN) {
        o.a = o.b = 0;
      } else if (len == NUTDB_PN_LEN_SPECIAL && gap == NUTDB_PN_GAP_EXT) {
        const NutdbNodeExt* e = find_ext(b, (uint32_t)(begin + j));
        if (!e) return false;
        o.kind = (uint8_t)(e->hdr & 255u);
        o.sub = (uint8_t)((e->hdr >> 8) & 255u);
        o.aux = (uint16_t)((e->hdr >> 16) & 1u);
        o.a = e->a;
        o.b = e->a + e->b;
        pos = o.b;
      } else {
        o.a = pos + gap;
        o.b = o.a + len;
        pos = o.b;
      }
    }
  }
  // pass 2: child counts and parent links
  auto subtree_start = [&](uint32_t r) -> uint32_t {
    if (out[r].kind >= NUTDB_NK_FIRST_INTERIOR) return out[r].a;
    return (out[r].kind == NUTDB_NK_IDENT && (out[r].aux & 1) && r > 0) ? r - 1 : r;
  };
  for (uint32_t j = 0; j < count; j++) {
    NutdbNode& o = out[j];
    if (o.kind >= NUTDB_NK_FIRST_INTERIOR) {
      uint32_t nchild = 0;
      int64_t r = (int64_t)j - 1;
      while (r >= (int64_t)o.a) {
        out[r].parent = j;
        nchild++;
        r = (int64_t)subtree_start((uint32_t)r) - 1;
      }
      o.b = nchild;
    } else if (o.kind == NUTDB_NK_IDENT && (o.aux & 1) && j > 0) {
      out[j - 1].parent = j;
    }
  }
  return true;
}

extern "C" {

size_t nutdb_fmt_debug(const NutdbBatch* batch, uint64_t i, const uint8_t* sql, size_t len, char* buf, size_t cap) {
  if (!batch || !batch->stmt || i >= batch->n_stmt) return 0;
  const NutdbStmt& s = batch->stmt[i];
  if (s.status != NUTDB_ST_OK || s.node_count == 0 || (!batch->node && !batch->pnode)) return deliver(std::string(), buf, cap);
  std::vector<NutdbNode> expanded;
  const NutdbNode* nodes;
  if (batch->node) {
    nodes = batch->node + s.node_begin;
  } else {  // the library's batches carry wire nodes: expand this statement's
    expanded.resize(s.node_count);
    if (!expand_stmt_nodes(batch, s.node_begin, s.node_count, expanded.data())) return deliver(std::string(), buf, cap);
    nodes = expanded.data();
  }
  H h{nodes, s.node_count, sql, len, std::string()};
  h.statement(s.node_count - 1);
  return deliver(h.o, buf, cap);
}

size_t nutdb_fmt_error(const NutdbBatch* batch, uint64_t i, const uint8_t* sql, size_t len, char* buf, size_t cap) {
  if (!batch || !batch->stmt || i >= batch->n_stmt) return 0;
  const NutdbError* e = find_error(batch, i);
  if (!e) return deliver(std::string(), buf, cap);
  std::string r;
  if (e->cls == NUTDB_ST_LEX_ERROR) r = lex_error_text(*e, sql, len);
  else if (e->cls == NUTDB_ST_SYNTAX_ERROR) r = syntax_error_text(*e, sql, len);
  else if (e->cls == NUTDB_ST_REFERENCE_PANIC) r = "the reference panics on this input (unreachable!() at literal.rs:63)";
  else r = "nesting exceeds the device parser's limits";
  return deliver(r, buf, cap);
}

int nutdb_batch_expand_stmts(const NutdbBatch* batch, NutdbStmt* out) {
  if (!batch || !out || (batch->n_stmt && !batch->wstmt && !batch->stmt)) return NUTDB_E_ARG;
  uint64_t node_begin = 0;
  for (uint64_t i = 0; i < batch->n_stmt; i++) {
    if (!batch->wstmt) {
      out[i] = batch->stmt[i];
      continue;
    }
    const uint64_t w = batch->wstmt[i];
    NutdbStmt& s = out[i];
    s.status = (uint32_t)(w & 15u);
    s.tok_begin = 0;
    s.tok_count = 0;
    s.node_begin = (uint32_t)node_begin;
    s.node_count = (uint32_t)((w >> 4) & 0x3FFFFFFFu);
    s.tok_used = (uint32_t)(w >> 34);
    node_begin += s.node_count;
  }
  return NUTDB_OK;
}

int nutdb_batch_expand_nodes(const NutdbBatch* batch, NutdbNode* out) {
  if (!batch || !out || (batch->n_node && !batch->pnode) || (batch->n_stmt && !batch->stmt && !batch->wstmt)) return NUTDB_E_ARG;
  std::vector<NutdbStmt> expanded;
  const NutdbStmt* stmts = batch->stmt;
  if (!stmts && batch->n_stmt) {
    expanded.resize(batch->n_stmt);
    nutdb_batch_expand_stmts(batch, expanded.data());
    stmts = expanded.data();
  }
  for (uint64_t i = 0; i < batch->n_stmt; i++) {
    const NutdbStmt& s = stmts[i];
    if (s.status == NUTDB_ST_OK && s.node_count && !expand_stmt_nodes(batch, s.node_begin, s.node_count, out + s.node_begin))
      return NUTDB_E_ARG;
  }
  return NUTDB_OK;
}

}  // extern "C"
