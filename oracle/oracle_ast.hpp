// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_lex.hpp header).
//
// C++ mirror of the reference AST (src/parser/ast/{mod,expr,item,query,alter}.rs), the literal
// helpers (src/parser/literal.rs) and the error types (src/parser/error.rs).  Fields marked
// "flat hint" do not exist in the reference: they only remember source spans / source order so
// that the tree can be serialised into the flat node format of include/nutdb_gpu.h.
//
// Third-party arithmetic: `bigdecimal = "0.3"` (Cargo.toml:18, Cargo.lock not committed, so the
// exact 0.3.x is unpinned and the crate is NOT vendored under /root/reference).  `Decimal` below
// restates its published behaviour from memory of bigdecimal 0.3.x: from_str = (BigInt of all
// digits, scale = number of fraction digits); PartialEq = numeric equality after rescaling;
// Display = plain decimal with exactly `scale` fraction digits; Debug = BigDecimal("<Display>").
// PARITY UNPINNED for BigDecimal values/equality/Debug (only call sites are pinned: literal.rs:7-14,
// mod.rs:1265,1283, simplify.rs via Literal: PartialEq).
#pragma once
#include <memory>
#include <optional>
#include <string>
#include <string_view>
#include <vector>

#include "oracle_lex.hpp"

namespace ora {

using sv = std::string_view;
typedef unsigned __int128 u128;

struct Decimal {
  bool neg = false;
  std::string mag = "0";  // magnitude digits, no leading zeros
  int64_t scale = 0;
  bool is_zero() const { return mag == "0"; }
  static Decimal from_str(sv s) {  // digits[.digits] as produced by the lexer
    Decimal d;
    size_t dot = s.find('.');
    std::string digits;
    if (dot == sv::npos) {
      digits = std::string(s);
      d.scale = 0;
    } else {
      digits = std::string(s.substr(0, dot)) + std::string(s.substr(dot + 1));
      d.scale = (int64_t)(s.size() - dot - 1);
    }
    size_t nz = digits.find_first_not_of('0');
    d.mag = nz == std::string::npos ? "0" : digits.substr(nz);
    return d;
  }
  Decimal negated() const {
    Decimal d = *this;
    if (!d.is_zero()) d.neg = !d.neg;
    return d;
  }
  bool operator==(const Decimal& o) const {
    if (is_zero() && o.is_zero()) return true;
    if (neg != o.neg) return false;
    std::string a = mag, b = o.mag;
    if (scale > o.scale) b += std::string((size_t)(scale - o.scale), '0');
    if (o.scale > scale) a += std::string((size_t)(o.scale - scale), '0');
    return a == b;
  }
  std::string display() const {
    std::string abs_int = mag, before, after;
    if (scale >= (int64_t)abs_int.size()) {
      after = std::string((size_t)(scale - (int64_t)abs_int.size()), '0') + abs_int;
      before = "0";
    } else {
      size_t loc = (size_t)((int64_t)abs_int.size() - scale);
      after = abs_int.substr(loc);
      before = abs_int.substr(0, loc);
    }
    std::string r = after.empty() ? before : before + "." + after;
    return (neg && !is_zero()) ? "-" + r : r;
  }
};

struct Query;

struct Expr {
  enum Kind : uint8_t { Identifier, QueryParameter, Literal, Collection, UnaryOp, BinaryOp, FnCall, Subquery };
  enum LitKind : uint8_t { LInteger, LFloat, LString, LBoolean, LInterval, LNull };
  Kind k = Literal;
  // Identifier: sub = 0 Word / 1 Wildcard; Literal: sub = LitKind; Collection: sub = CollectionType;
  // UnaryOp/BinaryOp: sub = operator ordinal; FnCall: sub = FnName ordinal
  uint8_t sub = 0;
  bool flag = false;  // Literal Integer: positive; Literal Boolean: value
  uint8_t unit = 0;   // Literal Interval: IntervalUnit
  sv s1;              // Identifier word / FnName::Others name / raw string literal
  sv s2;              // Identifier qualifier
  bool has_qual = false;
  u128 ival = 0;      // Literal Integer value / Interval value / QueryParameter index
  std::vector<Expr> kids;  // UnaryOp operand, BinaryOp left/right, Collection items, FnCall arguments
  std::unique_ptr<Query> q;            // Subquery
  std::unique_ptr<std::string> owned;  // Literal String (Cow::Owned)
  std::unique_ptr<Decimal> dec;        // Literal Float
  // flat hints
  Span span;           // source span of the literal token / word / index integer
  Span qspan;          // qualifier span
  uint8_t strkind = 0; // 0 raw 1 sq 2 dq
  bool hex = false;

  Expr() = default;
  Expr(Expr&&) noexcept = default;
  Expr& operator=(Expr&&) noexcept = default;

  sv str_value() const { return owned ? sv(*owned) : s1; }
  bool is_literal() const { return k == Literal; }
  bool is_bool() const { return k == Literal && sub == LBoolean; }
};

// Literal: PartialEq (derive, ast/item.rs:89)
inline bool literal_eq(const Expr& a, const Expr& b) {
  if (a.sub != b.sub) return false;
  switch (a.sub) {
    case Expr::LInteger: return a.ival == b.ival && a.flag == b.flag;
    case Expr::LFloat: return *a.dec == *b.dec;
    case Expr::LString: return a.str_value() == b.str_value();
    case Expr::LBoolean: return a.flag == b.flag;
    case Expr::LInterval: return a.ival == b.ival && a.unit == b.unit;
    default: return true;  // Null
  }
}

struct StringLit {  // Cow<'a, str> produced by must_parse_string_literal
  std::string value;
  Span span;
  uint8_t strkind = 0;
};

struct QueryExpr {
  Expr inner;
  std::optional<sv> alias;
};
struct QuerySource {
  enum Kind { TableFn, Table, SubqueryK } kind = Table;
  Expr fn;  // TableFn(FnCall)
  sv table;
  std::unique_ptr<Query> q;
  std::optional<sv> alias;
};
struct QueryCTE {
  std::unique_ptr<Query> subquery;
  sv alias;
};
struct OrderKey {
  QueryExpr expr;
  bool desc = false;
};
struct JoinClause {
  int typ = 0;
  QuerySource source;
  bool is_using = false;
  std::unique_ptr<Expr> on;
  std::vector<Expr> using_cols;  // Identifier exprs
};
struct LimitClause {
  size_t size = 0, offset = 0;
  bool with_ties = false;
  // flat hints
  int form = 0;
  Span s1, s2;
  bool hex1 = false, hex2 = false;
};
struct QueryBody {
  std::optional<std::vector<QueryCTE>> with;
  bool has_distinct = false;
  std::optional<std::vector<QueryExpr>> distinct_on;
  std::vector<QueryExpr> columns;
  std::optional<QuerySource> from;
  std::vector<JoinClause> joins;
  std::optional<Expr> where;
  std::optional<std::vector<QueryExpr>> group_by;
  std::optional<Expr> having;
  std::optional<std::vector<OrderKey>> order_by;
  std::optional<LimitClause> limit;
};
struct Query {
  bool is_union = false;
  std::unique_ptr<QueryBody> body;
  int typ = 0;
  std::unique_ptr<Query> left, right;
};

struct EnumBind {
  size_t id = 0;
  StringLit literal;
  // flat hints
  bool has_id = false;
  Span idspan;
  bool idhex = false;
};
struct DataType {
  bool compound = false;
  int id = 0;  // scalar index 0..25 or compound kind 0..5
  size_t param = 0;
  std::vector<DataType> inner;  // Map holds (value, key) like the reference (mod.rs:1780)
  std::vector<EnumBind> binds;
  // flat hints
  bool has_param = false;
  Span pspan;
  bool phex = false;
};
struct ColumnDefinition {
  sv name;
  DataType typ;
  std::optional<Expr> default_;
  std::optional<StringLit> comment;
  std::vector<int> attr_order;  // flat hint: 0 default, 1 comment
};
struct ConstraintDefinition {
  sv name;
  Expr check;
};
struct IndexDefinition {
  sv name;
  Expr indexer;  // FnCall
};
struct TableDefinition {
  sv name;
  std::vector<ColumnDefinition> columns;
  std::vector<ConstraintDefinition> constraints;
  std::vector<IndexDefinition> indexes;
  std::optional<std::vector<Expr>> primary_key, order_by;
  std::optional<Expr> partition_by;
  std::optional<StringLit> comment;
  std::vector<std::pair<int, size_t>> item_order;  // flat hint: (0 column | 1 index | 2 constraint, idx)
  std::vector<int> attr_order;                     // flat hint: 0 pk 1 order 2 partition 3 comment
};
struct ViewDefinition {
  sv name;
  sv strategy;
  std::optional<std::vector<Expr>> primary_key, order_by;
  std::optional<Expr> partition_by;
  std::unique_ptr<Query> query;
  std::optional<StringLit> comment;
  std::vector<int> attr_order;  // flat hint: 0 update 1 pk 2 order 3 partition 4 comment
};

struct Statement {
  enum Kind { Select, Insert, Explain, Alter, Create, Describe, Drop, Truncate, Optimize, Set } k = Select;
  std::unique_ptr<Query> query;  // Select / Explain / Insert subquery
  // Insert
  sv table_name;
  std::optional<std::vector<sv>> column_list;
  int insert_kind = 0;  // 0 rows 1 subquery 2 fncall
  size_t column_size = 0;
  std::vector<Expr> rows_data;
  Expr insert_fn;
  // Alter
  int alter_action = 0;  // 0 add 1 drop 2 rename
  int entity_kind = 0;   // add: 0 column 1 index 2 constraint (source kw order); drop: +3 partition; rename: +3 table
  bool flag = false;     // if_not_exists / if_exists
  ColumnDefinition col;
  IndexDefinition idx;
  ConstraintDefinition con;
  int position = 2;  // 0 first 1 after 2 last
  sv after_name;
  sv entity_name;
  StringLit partition;
  sv new_name;
  // Create
  bool is_view = false;
  TableDefinition table;
  ViewDefinition view;
  // Describe / Drop / Truncate
  int entity = 0;  // describe: 0 table 1 view 2 database; drop/truncate: 0 table 1 view
  sv name;
  // Optimize
  std::optional<Expr> partition_key;
  // Set
  sv config_name;
  Expr value;
};

// ------------------------------------------------------------------------------------------
// errors (src/parser/error.rs)
// ------------------------------------------------------------------------------------------
struct SyntaxError {
  int variant = 0;  // NUTDB_SE_*
  std::vector<TT> expected_types;
  TT actual_type = 0;
  std::vector<std::string> expected_kw;
  std::string actual_kw;
  std::string msg;
  bool has_pos = false;
  Position pos;
  size_t byte_pos = 0;
  std::string raw;  // hex / raw literal text
  std::string this_, that;
  // record fields (include/nutdb_gpu.h NutdbError.a/b/c)
  uint32_t a = 0, b = 0, c = 0;
};

struct ReferencePanic {};  // the reference hits unreachable!() (literal.rs:63) -- it would panic, not return

struct ParseErr {
  bool is_lex = false;
  TokenizeError lex;
  SyntaxError syn;
};

// ------------------------------------------------------------------------------------------
// literal.rs
// ------------------------------------------------------------------------------------------
// u128::from_str / from_str_radix on lexer-produced digit strings: only Empty or PosOverflow can fail.
inline bool parse_uint(sv s, int radix, u128 max, u128& out) {
  if (s.empty()) return false;
  u128 v = 0;
  for (char ch : s) {
    unsigned d;
    if (ch >= '0' && ch <= '9') d = (unsigned)(ch - '0');
    else if (ch >= 'a' && ch <= 'f') d = (unsigned)(ch - 'a' + 10);
    else if (ch >= 'A' && ch <= 'F') d = (unsigned)(ch - 'A' + 10);
    else return false;
    if (d >= (unsigned)radix) return false;
    if (v > (max - d) / (unsigned)radix) return false;
    v = v * (unsigned)radix + d;
  }
  out = v;
  return true;
}

// literal.rs:36-102; `quote` is '\'' or '"'.  hex_span (relative to raw) reports the collected hex.
inline bool unescape_string(sv raw, char32_t quote, std::string& res, std::string& bad_hex, Span& hex_span) {
  res.clear();
  res.reserve(raw.size());
  // iterate chars
  size_t i = 0;
  auto next_char = [&](char32_t& out, size_t& at) -> bool {
    if (i >= raw.size()) return false;
    at = i;
    unsigned char x = (unsigned char)raw[i];
    size_t n = x < 0x80 ? 1 : (x >= 0xF0 ? 4 : (x >= 0xE0 ? 3 : 2));
    char32_t c;
    if (n == 1) c = x;
    else if (n == 2) c = ((x & 0x1F) << 6) | (raw[i + 1] & 0x3F);
    else if (n == 3) c = ((x & 0x0F) << 12) | ((raw[i + 1] & 0x3F) << 6) | (raw[i + 2] & 0x3F);
    else c = ((x & 0x07) << 18) | ((raw[i + 1] & 0x3F) << 12) | ((raw[i + 2] & 0x3F) << 6) | (raw[i + 3] & 0x3F);
    i += n;
    out = c;
    return true;
  };
  char32_t ch;
  size_t at;
  while (next_char(ch, at)) {
    if (ch == quote) {
      char32_t dummy;
      size_t d;
      next_char(dummy, d);  // chars.next()
      res += encode_utf8(quote);
    } else if (ch == '\\') {
      char32_t nc;
      size_t nat;
      if (!next_char(nc, nat)) throw ReferencePanic{};  // unreachable!() literal.rs:63: reachable through `\\u` eating a char
      if (nc == 'n') res.push_back('\n');
      else if (nc == 'r') res.push_back('\r');
      else if (nc == 't') res.push_back('\t');
      else if (nc == 'u') {
        char32_t b;
        size_t bat;
        if (next_char(b, bat) && b == '{') {
          // chars.by_ref().take_while(|&ch| ch != '}').collect()
          size_t hs = i;
          std::string hex;
          size_t he = i;
          char32_t hc;
          size_t hat;
          while (next_char(hc, hat)) {
            if (hc == '}') break;
            hex += encode_utf8(hc);
            he = i;
          }
          // u32::from_str_radix(&hex, 16): accepts a leading '+', rejects empty / lone sign / overflow
          bool ok = true;
          uint64_t v = 0;
          sv h = hex;
          if (!h.empty() && h[0] == '+') h.remove_prefix(1);
          if (h.empty()) ok = false;
          for (char cc : h) {
            unsigned d;
            if (cc >= '0' && cc <= '9') d = (unsigned)(cc - '0');
            else if (cc >= 'a' && cc <= 'f') d = (unsigned)(cc - 'a' + 10);
            else if (cc >= 'A' && cc <= 'F') d = (unsigned)(cc - 'A' + 10);
            else { ok = false; break; }
            v = v * 16 + d;
            if (v > 0xFFFFFFFFull) { ok = false; break; }
          }
          // char::from_u32
          if (ok && (v > 0x10FFFF || (v >= 0xD800 && v <= 0xDFFF))) ok = false;
          if (!ok) {
            bad_hex = hex;
            hex_span = Span{hs, he};
            return false;
          }
          res += encode_utf8((char32_t)v);
        } else {
          res.push_back('u');  // and the char after 'u' is dropped (literal.rs:70,89-91)
        }
      } else {
        res += encode_utf8(nc);
      }
    } else {
      res += encode_utf8(ch);
    }
  }
  return true;
}

}  // namespace ora
